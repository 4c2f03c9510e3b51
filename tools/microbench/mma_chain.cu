// tools/microbench/mma_chain.cu -- what does one MLP layer cost on an SM, and how does it scale with batches in flight?
//
// NOT RUN YET (written at the end of round 1, after the GPU budget was spent): run it under `timeout`, it spins on
// mbarriers and a protocol slip would hang it.
//
// The head kernel's MLP is a serial chain per 128-row batch:  MMA (128x128x128, fp16, SS operands, accumulator in TMEM) ->
// tcgen05.commit -> all rows read the accumulator (tcgen05.ld), ReLU, write the next layer's fp16 operand tiles -> next MMA.
// In v1 one CTA runs one such chain at a time (~1.2 K cycles MMA phase + 1.0-1.6 K epilogue per layer, two CTAs per SM).
// docs/HEAD_V2_PLAN.md wants several batch slots in flight inside one CTA; this probe measures exactly that skeleton:
//
//   warp 0            issue warp: for every layer, for every slot: wait a_full[slot] -> 8 x tcgen05.mma -> commit acc_full[slot]
//   warps 4+4s..7+4s  epilogue warps of slot s (one row per thread): wait acc_full[s] -> 128 columns TMEM -> ReLU -> fp16 ->
//                     SW128 operand tiles of slot s -> fence.proxy.async -> arrive a_full[s]
//
// and reports cycles per layer per slot and layers per kilocycle per SM for 1, 2 and 3 slots (weights resident: two 16 KB
// tiles; the streaming of weights is measured separately by the real kernel).  `light` epilogue = read TMEM only (no
// conversion / operand stores) to split the epilogue cost.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I genefaceplusplus_b200/csrc -o tools/microbench/mma_chain.bin tools/microbench/mma_chain.cu
//   timeout 30 tools/microbench/mma_chain.bin
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "tc.cuh"

using namespace gfpp::tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int TILE = 16384;   // 128 rows x 64 k x 2 B, SW128

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}

template <int NSLOT>
struct alignas(1024) Smem {
    unsigned char a[NSLOT][2][TILE];
    unsigned char w[2][TILE];
    unsigned long long a_full[NSLOT], acc_full[NSLOT];
    uint32_t tmem_base;
};

template <int NSLOT>
__global__ void __launch_bounds__(128 + 128 * NSLOT, 1) k_chain(int layers, int light, unsigned long long *cycles, float *sink) {
    extern __shared__ __align__(1024) unsigned char raw_[];
    unsigned char *raw = raw_ + ((1024u - (smem_u32(raw_) & 1023u)) & 1023u);
    Smem<NSLOT> &s = *reinterpret_cast<Smem<NSLOT> *>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t TMEM_COLS = NSLOT <= 1 ? 128 : (NSLOT == 2 ? 256 : 512);

    if (warp == 0) tmem_alloc(&s.tmem_base, TMEM_COLS);
    if (tid == 32) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(&s.a_full[i], 128); mbar_init(&s.acc_full[i], 1); }
        mbar_fence_init();
    }
    // weights: small pseudo-random fp16 values, written straight in the swizzled layout (any values do for timing)
    for (int i = tid; i < 2 * TILE / 4; i += blockDim.x) {
        const uint32_t h = (uint32_t)i * 2654435761u;
        reinterpret_cast<uint32_t *>(&s.w[0][0])[i] = pack2<false>(((h >> 8) & 255) / 2048.0f - 0.06f, ((h >> 16) & 255) / 2048.0f - 0.06f);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s.tmem_base;
    const uint32_t idesc = make_idesc(0, 128);
    const long long t0 = clock64();

    if (warp == 0) {
        // ---------------- issue warp ----------------
        for (int l = 0; l < layers; ++l) {
#pragma unroll
            for (int sl = 0; sl < NSLOT; ++sl) {
                mbar_wait(&s.a_full[sl], l & 1);
                fence_after_sync();
                if (elect_one()) {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = desc_sw128(smem_u32(s.a[sl][c]) + k * 32u), dw = desc_sw128(smem_u32(s.w[c]) + k * 32u);
                            if (c == 0 && k == 0) mma_f16_s<false>(tmem + sl * 128u, da, dw, idesc);
                            else mma_f16_s<true>(tmem + sl * 128u, da, dw, idesc);
                        }
                    mma_commit(&s.acc_full[sl]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue warps: slot = (warp - 4) / 4, TMEM lane quarter = warp % 4 ----------------
        const int sl = (warp - 4) >> 2, row = (tid - 128) & 127;
        const uint32_t taddr = tmem + ((uint32_t)(warp & 3) * 32u << 16) + (uint32_t)sl * 128u;
        float acc_keep = 0.f;
        {   // initial operand: row-dependent small values
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.01f * (float)((row + i) & 15);
            for (int c = 0; c < 16; ++c) store_chunk<false, false>(s.a[sl][c >> 3], nullptr, sw128_off(row, c & 7), v);
            fence_async_smem();
            mbar_arrive(&s.a_full[sl]);
        }
        for (int l = 0; l < layers; ++l) {
            mbar_wait(&s.acc_full[sl], l & 1);
            fence_after_sync();
#pragma unroll 1
            for (int p = 0; p < 4; ++p) {
                float v[2][16];
                tmem_ld16(taddr + p * 32, v[0]);
                tmem_ld16(taddr + p * 32 + 16, v[1]);
                wait_ld();
#pragma unroll
                for (int q = 0; q < 2; ++q) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[q][i] = fmaxf(v[q][i], 0.f) * 0.5f + 0.01f;
                    if (light) {
                        acc_keep += v[q][0] + v[q][15];
                    } else {
                        const int col = p * 32 + q * 16;   // 0..127 -> tile col>>6, chunk (col & 63) >> 3
                        store_chunk<false, false>(s.a[sl][col >> 6], nullptr, sw128_off(row, (col & 63) >> 3), &v[q][0]);
                        store_chunk<false, false>(s.a[sl][col >> 6], nullptr, sw128_off(row, ((col & 63) >> 3) + 1), &v[q][8]);
                    }
                }
            }
            fence_async_smem();
            fence_before_sync();
            mbar_arrive(&s.a_full[sl]);
        }
        if (acc_keep == 12345.678f) sink[0] = acc_keep;
    }
    const long long t1 = clock64();
    if (tid == 128) atomicAdd(cycles, (unsigned long long)(t1 - t0));   // slot 0's first epilogue thread: the whole chain
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

template <int NSLOT>
static void run(int sms, int light, unsigned long long *d_cyc, float *d_sink) {
    const int layers = 600;
    const size_t smem = sizeof(Smem<NSLOT>) + 1024;
    CK(cudaFuncSetAttribute(k_chain<NSLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaMemset(d_cyc, 0, 8));
    k_chain<NSLOT><<<sms, 128 + 128 * NSLOT, smem>>>(20, light, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(d_cyc, 0, 8));
    k_chain<NSLOT><<<sms, 128 + 128 * NSLOT, smem>>>(layers, light, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    unsigned long long cyc = 0;
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    const double per_layer = (double)cyc / sms / layers;   // cycles between consecutive layers of ONE slot
    printf("slots=%d %s epilogue: %7.0f cycles per layer per slot, %6.3f layer-batches per kilocycle per SM (tensor floor 512 cyc/layer)\n", NSLOT,
           light ? "light" : "full ", per_layer, NSLOT * 1000.0 / per_layer);
    fflush(stdout);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    unsigned long long *d_cyc; float *d_sink;
    CK(cudaMalloc(&d_cyc, 8)); CK(cudaMalloc(&d_sink, 4));
    printf("%s, %d SMs; layer = 128x128x128 fp16 MMA (8 x tcgen05.mma) + 128-column epilogue; one CTA per SM\n", prop.name, prop.multiProcessorCount);
    for (int light = 1; light >= 0; --light) {
        run<1>(prop.multiProcessorCount, light, d_cyc, d_sink);
        run<2>(prop.multiProcessorCount, light, d_cyc, d_sink);
        run<3>(prop.multiProcessorCount, light, d_cyc, d_sink);
    }
    return 0;
}
