// head_tc_kernel.cu -- fused persistent head renderer, tensor-core variant (tcgen05 + TMEM + TMA bulk copies).
//
// Same ray-slot machinery, marching, gather and compositing as head_kernel.cu; only the dense tiny-MLP GEMMs differ:
// every 128-row batch runs its eight layers as tcgen05.mma (kind::f16, M=128, N in {128,144,16}) with
//   * A (activations) written by the CTA's threads straight into UMMA K-major SW128 tiles in shared memory
//     (gathered grid features, conditioning, previous layer's epilogue), 16-bit, optionally hi/lo split;
//   * W streamed as pre-swizzled 16-bit tiles through a 3-stage ring with cp.async.bulk + mbarrier;
//   * D (fp32 accumulators) in TMEM, read back with tcgen05.ld by all 256 threads (2 threads per row) for the
//     ReLU / tanh / exp / sigmoid epilogues;
//   * position-grid features parked in 32 spare TMEM columns between the ambient and sigma nets.
// One thread issues all MMAs; completion is signalled with tcgen05.commit on an mbarrier.
//
// Precision modes (tc::Precision): FP16_X1 (what the reference itself runs under autocast, fp32 accumulate),
// BF16_X1, BF16_X3 (hi/lo split, 3 MMAs per k-step: ~16 mantissa bits).  See DESIGN.md "precision".
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "gather8.cuh"
#include "head_common.cuh"
#include "head_kernel.cuh"
#include "launch.cuh"
#include "tc.cuh"

namespace gfpp {

using namespace tc;
using namespace headc;

namespace {

constexpr int TM = HEAD_TM;
constexpr int NT = HEAD_NT;
constexpr int A_TILE = 16384;        // 128 rows x 64 k x 2 B (SW128)
constexpr int A_K16 = 4096;          // 128 rows x 16 k x 2 B (no swizzle)
constexpr int W_HALF = 18432;        // up to 144 rows x 64 k x 2 B
constexpr int W_NSTAGE = 3;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t TMEM_P = 160;     // 32 columns of parked position features

// the lo (split) operand images exist only in the bf16x3 instantiation; the single-pass kernels stay under 113 KB of
// shared memory and 128 registers so that TWO CTAs are resident per SM (their barrier / latency stalls overlap)
template <bool SPLIT>
struct SmemTC {
    // SW128 tiles must start on 1024-byte boundaries (the swizzle is a function of the shared-memory address bits)
    alignas(1024) unsigned char a_hi[2][A_TILE];
    alignas(1024) unsigned char w_hi[W_NSTAGE][W_HALF];
    alignas(SPLIT ? 1024 : 16) unsigned char a_lo[SPLIT ? 2 : 1][SPLIT ? A_TILE : 16];
    alignas(SPLIT ? 1024 : 16) unsigned char w_lo[SPLIT ? W_NSTAGE : 1][SPLIT ? W_HALF : 16];
    alignas(256) unsigned char s_hi[A_K16];
    alignas(SPLIT ? 256 : 16) unsigned char s_lo[SPLIT ? A_K16 : 16];
    float bias[128];                  // color L0 bias (individual code folded)
    float nw[6 * 128];                // fp32 rows of the two 3-wide output layers: ambient out (0-2), color out (3-5)
    float part[2 * 3 * TM];           // their partial dot products, one per column half
    float sx[TM], sy[TM], sz[TM];
    float amb[3 * TM];
    float sig[TM];
    float rgb[3 * TM];
    int frame[TM];
    int valid[TM];
    unsigned long long bar_full[W_NSTAGE];
    unsigned long long bar_acc;
    uint32_t tmem_base;
    uint32_t coarse[HEAD_COARSE_WORDS];
    uint4 lvl[2][2 * GFPP_MAX_LEVELS];   // per-level grid metadata (position, ambient) for lookup8o
    int next, end, done;
};

// The weight stream is static: batch b consumes tiles b*12 + (0..11) in the order of kHeadTcChunks, and 12 is a multiple
// of the ring depth, so tile id c always lives in stage c % W_NSTAGE and completes the mbarrier phase (c / W_NSTAGE) & 1.
// Everything the issuing warp needs is therefore a compile-time constant or a CTA-uniform address: the MMA / bulk-copy
// operands sit in uniform registers and one elected lane of warp 0 issues them back to back.  (A lone `tid == 0` branch
// makes the compiler wrap every tcgen05.mma in a register->uniform-register waterfall loop: ~250 cycles per MMA.)
static_assert(HEAD_TC_NCHUNK % W_NSTAGE == 0, "the static stage/parity schedule needs the ring depth to divide the tiles per batch");

template <bool SPLIT, int ID>
__device__ __forceinline__ void issue_load(const HeadTcArgs &t, SmemTC<SPLIT> &s) {
    constexpr int id = ID % HEAD_TC_NCHUNK, stage = id % W_NSTAGE;
    constexpr uint32_t bytes = (uint32_t)head_tc_chunk_bytes(id), off = (uint32_t)head_tc_chunk_off(id);
    mbar_expect_tx(&s.bar_full[stage], SPLIT ? 2 * bytes : bytes);
    bulk_g2s(s.w_hi[stage], t.w_hi + off, bytes, &s.bar_full[stage]);
    if (SPLIT) bulk_g2s(s.w_lo[SPLIT ? stage : 0], t.w_lo + off, bytes, &s.bar_full[stage]);
}
template <bool SPLIT, int ID, int N>
__device__ __forceinline__ void issue_loads(const HeadTcArgs &t, SmemTC<SPLIT> &s) {
    if constexpr (N > 0) {
        issue_load<SPLIT, ID>(t, s);
        issue_loads<SPLIT, ID + 1, N - 1>(t, s);
    }
}

template <bool SPLIT, int ID, bool ACC0>
__device__ __forceinline__ void issue_tile_mmas(SmemTC<SPLIT> &s, int a_tile, uint32_t d_tmem, uint32_t idesc) {
    constexpr int stage = ID % W_NSTAGE, ksteps = head_tc_chunk_ksteps(ID);
    constexpr bool k16 = kHeadTcChunks[ID].k16 != 0;
    const uint32_t ah = k16 ? smem_u32(s.s_hi) : smem_u32(s.a_hi[a_tile]);
    const uint32_t al = k16 ? smem_u32(s.s_lo) : smem_u32(s.a_lo[SPLIT ? a_tile : 0]);
    const uint32_t wh = smem_u32(s.w_hi[stage]), wl = smem_u32(s.w_lo[SPLIT ? stage : 0]);
#pragma unroll
    for (int k = 0; k < ksteps; ++k) {
        const uint32_t adv = (uint32_t)k * 32u;   // 16 elements x 2 B inside the 128-byte swizzle row
        const uint64_t dah = k16 ? desc_k16(ah) : desc_sw128(ah + adv), dwh = k16 ? desc_k16(wh) : desc_sw128(wh + adv);
        if (SPLIT) {
            const uint64_t dal = k16 ? desc_k16(al) : desc_sw128(al + adv), dwl = k16 ? desc_k16(wl) : desc_sw128(wl + adv);
            if (ACC0 || k > 0) mma_f16_s<true>(d_tmem, dal, dwh, idesc);   // small terms first
            else mma_f16_s<false>(d_tmem, dal, dwh, idesc);
            mma_f16_s<true>(d_tmem, dah, dwl, idesc);
            mma_f16_s<true>(d_tmem, dah, dwh, idesc);
        } else {
            if (ACC0 || k > 0) mma_f16_s<true>(d_tmem, dah, dwh, idesc);
            else mma_f16_s<false>(d_tmem, dah, dwh, idesc);
        }
    }
}
template <bool SPLIT, int ID, int C, int NCH>
__device__ __forceinline__ void issue_tiles(SmemTC<SPLIT> &s, uint32_t d_tmem, uint32_t idesc) {
    if constexpr (C < NCH) {
        constexpr int stage = ID % W_NSTAGE;
        mbar_wait(&s.bar_full[stage], (ID / W_NSTAGE) & 1);
        fence_after_sync();
        if (elect_one()) issue_tile_mmas<SPLIT, ID, (C > 0)>(s, C, d_tmem, idesc);
        __syncwarp();
        issue_tiles<SPLIT, ID + 1, C + 1, NCH>(s, d_tmem, idesc);
    }
}

// Warp 0 (all lanes, converged): issue one layer = tiles ID0 .. ID0+NCH-1 against A tiles 0,1,.. (the K16 SH tile for the
// k16 chunk), then commit to the accumulator barrier.
template <bool SPLIT, int ID0, int NCH>
__device__ __forceinline__ void issue_layer(SmemTC<SPLIT> &s, uint32_t d_tmem, uint32_t idesc) {
    fence_after_sync();
    issue_tiles<SPLIT, ID0, 0, NCH>(s, d_tmem, idesc);
    if (elect_one()) mma_commit(&s.bar_acc);
    __syncwarp();
}

// All threads: wait for the accumulator; warp 0 then refills the weight stages the layer has released.
template <bool SPLIT, int ID0, int NCH>
__device__ __forceinline__ void wait_acc(const HeadTcArgs &t, SmemTC<SPLIT> &s, uint32_t &acc_uses, int warp) {
    mbar_wait(&s.bar_acc, acc_uses & 1);
    acc_uses += 1;
    fence_after_sync();
    if (warp == 0) {
        if (elect_one()) issue_loads<SPLIT, ID0 + W_NSTAGE, NCH>(t, s);
        __syncwarp();
    }
}

// Epilogue of a 128-wide layer: thread (row, half) reads 64 accumulator columns (two passes of 32), applies bias/ReLU and
//   * NARROW == false: writes them as the next layer's A operand (k = column) into tile `half`;
//   * NARROW == true : the layer feeds only a 3-wide output layer (ambient coordinates / rgb), so the activations never go
//     back to shared memory: the thread accumulates its half of the three dot products in fp32 (weights `nw`, [3][128])
//     and stores the partials in part[half][o][row] -- no fp16 round trip, no extra MMA, no extra barrier.
// Not inlined: it is called six times per batch and the kernel must stay instruction-cache friendly.
template <bool BF16, bool SPLIT, bool RELU, bool BIAS, bool NARROW>
__device__ __noinline__ void epilogue_wide(unsigned char *a_hi, unsigned char *a_lo, uint32_t tmem, int tid, const float *bias,
                                           const float *nw, float *part) {
    const int row = tid & 127, half = tid >> 7;
    const uint32_t lane_base = (uint32_t)((tid >> 5) & 3) * 32u;
    const uint32_t taddr = tmem + (lane_base << 16) + (uint32_t)half * 64u;
    unsigned char *hi = a_hi + half * A_TILE, *lo = a_lo + (SPLIT ? half * A_TILE : 0);
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll 1
    for (int p = 0; p < 2; ++p) {
        float v[2][16];
        tmem_ld16(taddr + p * 32, v[0]);
        tmem_ld16(taddr + p * 32 + 16, v[1]);
        wait_ld();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int col = half * 64 + p * 32 + q * 16 + i;
                float x = v[q][i];
                if (BIAS) x += bias[col];
                if (RELU) x = fmaxf(x, 0.f);
                v[q][i] = x;
                if (NARROW) {
                    d0 = fmaf(x, nw[col], d0);
                    d1 = fmaf(x, nw[128 + col], d1);
                    d2 = fmaf(x, nw[256 + col], d2);
                }
            }
            if (!NARROW) {
                store_chunk<BF16, SPLIT>(hi, lo, sw128_off(row, 4 * p + 2 * q), &v[q][0]);
                store_chunk<BF16, SPLIT>(hi, lo, sw128_off(row, 4 * p + 2 * q + 1), &v[q][8]);
            }
        }
    }
    if (NARROW) {
        part[(half * 3 + 0) * TM + row] = d0;
        part[(half * 3 + 1) * TM + row] = d1;
        part[(half * 3 + 2) * TM + row] = d2;
    }
}

// out of line on purpose (I-cache, register budget of the 2-CTA kernel)
__device__ __noinline__ void lookup8o(const uint4 *lvl, float align_off, bool smoothstep, const uint4 *__restrict__ octs, float u, float v,
                                      float w, float (&f)[16]) {
    lookup8<OCT_F16>(lvl, align_off, smoothstep, octs, u, v, w, f);
}

// Four consecutive levels of a grid -> 8 features (one 16-byte operand chunk): the fp32 quad layout (bf16x3 / bf16 modes) or
// the reference layout.  Out of line on purpose (I-cache).
__device__ __noinline__ void lookup4(const GridMeta &gm, const float2 *__restrict__ table, const float4 *__restrict__ quads, int l0,
                                     float u, float v, float w, float (&f)[8]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 g2 = (gm.dim == 3) ? (quads ? grid_lookup3q(gm, quads, l0 + j, u, v, w) : grid_lookup3(gm, table, l0 + j, u, v, w))
                                        : grid_lookup2(gm, table, l0 + j, u, v);
        f[2 * j] = g2.x;
        f[2 * j + 1] = g2.y;
    }
}

}  // namespace

template <bool BF16, bool SPLIT>
__global__ void __launch_bounds__(HEAD_NT, SPLIT ? 1 : 2) k_head_tc(const __grid_constant__ HeadArgs a, const __grid_constant__ HeadTcArgs t) {
    extern __shared__ __align__(1024) unsigned char smem_raw_[];
    unsigned char *smem_raw = smem_raw_ + ((1024u - (smem_u32(smem_raw_) & 1023u)) & 1023u);
    SmemTC<SPLIT> &s = *reinterpret_cast<SmemTC<SPLIT> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- one-time setup ----
    MarchConst mc = a.mc;
    float occ_lo[3] = {0.f, 0.f, 0.f}, occ_hi[3] = {0.f, 0.f, 0.f};
    const bool have_box = setup_occupancy(a, mc, occ_lo, occ_hi);
    install_coarse(a, s, mc, tid, NT);
    if (warp == 0) tmem_alloc(&s.tmem_base, TMEM_COLS);
    if (tid == 32) {
        for (int i = 0; i < W_NSTAGE; ++i) mbar_init(&s.bar_full[i], 1);
        mbar_init(&s.bar_acc, 1);
        mbar_fence_init();
        s.next = 0; s.end = 0; s.done = 0;
    }
    for (int i = tid; i < 3 * 128; i += NT) {   // a.narrow rows: 0-2 ambient out, 3 sigma, 4-6 color out, 7 color-L0 bias
        s.nw[i] = a.narrow[i];
        s.nw[3 * 128 + i] = a.narrow[4 * 128 + i];
    }
    if (tid < 128) s.bias[tid] = a.narrow[7 * 128 + tid];
    stage_level_meta(a.pos_gm, nullptr, s.lvl[0], tid);
    stage_level_meta(a.amb_gm, nullptr, s.lvl[1], tid);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s.tmem_base;
    if (warp == 0) {
        if (elect_one()) issue_loads<SPLIT, 0, W_NSTAGE>(t, s);
        __syncwarp();
    }
    uint32_t acc_uses = 0;   // completed accumulator hand-offs (uniform across the CTA)
    const uint32_t idesc128 = make_idesc(BF16 ? 1 : 0, 128), idesc144 = make_idesc(BF16 ? 1 : 0, 144);
    const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;

    const int total = (a.pass == 1) ? *a.n_hits : *a.n_survivors;
    Slot sl;
    sl.active = false;
    sl.gid = 0; sl.frame = 0; sl.nsamp = 0; sl.cap = 0;

    long long ph_last = clock64();
#define PH(i)                                                                                   \
    if (a.phase_cycles && tid == 0) {                                                           \
        const long long now_ = clock64();                                                       \
        atomicAdd(a.phase_cycles + (i), (unsigned long long)(now_ - ph_last));                  \
        ph_last = now_;                                                                         \
    }
    for (;;) {
        // ================= refill dead slots from the global cursor, publish the batch (head_common.cuh) =================
        const int n_valid = refill_and_publish(a, s, sl, mc, have_box, occ_lo, occ_hi, total, tid);
        if (n_valid < 0) break;
        if (n_valid == 0) continue;
        PH(0)   // refill + publish

        const int slot = tid & (TM - 1), lg = tid >> 7;
        const bool v = s.valid[slot] != 0;
        // ---- position grid -> A tile0 k[0,32) and TMEM park; conditioning -> tile0 k[32,64), tile1 k[0,32) ----
        {
            float u = 0.f, vv = 0.f, w = 0.f;
            if (v) {
                const float inv2b = 2.0f * mc.bound;
                u = __fdiv_rn(__fadd_rn(s.sx[slot], mc.bound), inv2b);
                vv = __fdiv_rn(__fadd_rn(s.sy[slot], mc.bound), inv2b);
                w = __fdiv_rn(__fadd_rn(s.sz[slot], mc.bound), inv2b);
            }
            if (a.pos_octs) {   // thread (slot, lg) owns levels lg*8 .. lg*8+7 = operand chunks 2lg, 2lg+1
                float f[16];
                if (v) lookup8o(&s.lvl[0][lg * 16], a.pos_gm.align_off, a.pos_gm.interp == 1, a.pos_octs, u, vv, w, f);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = 0.f;
                }
                store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 2 * lg), &f[0]);
                store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 2 * lg + 1), &f[8]);
                tmem_st16(tmem + (lane_base << 16) + TMEM_P + (uint32_t)(lg * 16), f);
            } else {
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    float f[8];
                    if (v) lookup4(a.pos_gm, a.pos_tab, a.pos_quads, lg * 8 + c * 4, u, vv, w, f);
                    else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) f[i] = 0.f;
                    }
                    store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 2 * lg + c), f);
                    tmem_st8(tmem + (lane_base << 16) + TMEM_P + (uint32_t)(lg * 16 + c * 8), f);
                }
            }
            // cond values [lg*32, lg*32+32) sit at k = 32 + lg*32 + i: lg 0 -> tile0 chunks 4..7, lg 1 -> tile1 chunks 0..3
            const float4 *cf = reinterpret_cast<const float4 *>(a.cond_feat + (size_t)s.frame[slot] * 64 + lg * 32);
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                float c8[8];
                if (v) {
                    const float4 q0 = __ldg(cf + 2 * j), q1 = __ldg(cf + 2 * j + 1);
                    c8[0] = q0.x; c8[1] = q0.y; c8[2] = q0.z; c8[3] = q0.w; c8[4] = q1.x; c8[5] = q1.y; c8[6] = q1.z; c8[7] = q1.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) c8[i] = 0.f;
                }
                store_chunk<BF16, SPLIT>(s.a_hi[lg], s.a_lo[SPLIT ? lg : 0], sw128_off(slot, (lg == 0 ? 4 : 0) + j), c8);
            }
            wait_st();
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        PH(1)   // position gather + cond + park

        // ---- ambient net 96 -> 128 -> 128 -> 3 ----
        if (warp == 0) issue_layer<SPLIT, 0, 2>(s, tmem, idesc128);
        wait_acc<SPLIT, 0, 2>(t, s, acc_uses, warp);
        PH(2)   // MMA issue + wait (ambient L0)
        epilogue_wide<BF16, SPLIT, true, false, false>(&s.a_hi[0][0], &s.a_lo[0][0], tmem, tid, nullptr, nullptr, nullptr);
        fence_async_smem(); fence_before_sync(); __syncthreads();
        PH(3)   // epilogue (ambient L0) + barrier
        if (warp == 0) issue_layer<SPLIT, 2, 2>(s, tmem, idesc128);
        wait_acc<SPLIT, 2, 2>(t, s, acc_uses, warp);
        PH(4)   // MMA (ambient L1)
        // ambient L1 epilogue fused with the 3-wide ambient output layer (fp32 dot products straight from the accumulators)
        epilogue_wide<BF16, SPLIT, true, false, true>(&s.a_hi[0][0], &s.a_lo[0][0], tmem, tid, nullptr, s.nw, s.part);
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        PH(5)   // epilogue (ambient L1) + partial dots
        if (tid < TM) {
            s.amb[tid] = tanhf(s.part[0 * TM + tid] + s.part[3 * TM + tid]);
            s.amb[TM + tid] = tanhf(s.part[1 * TM + tid] + s.part[4 * TM + tid]);
            s.amb[2 * TM + tid] = tanhf(s.part[2 * TM + tid] + s.part[5 * TM + tid]);
        }
        __syncthreads();
        PH(6)   // narrow ambient out: MMA + tanh
        // ---- sigma-net input: tile0 k[0,32) <- parked position features, k[32,64) <- ambient grid ----
        {
            float u = 0.f, vv = 0.f, w = 0.f;
            if (v) {   // GridEncoder.forward with bound = 1: (x + 1) / 2
                u = __fdiv_rn(__fadd_rn(s.amb[slot], 1.0f), 2.0f);
                vv = __fdiv_rn(__fadd_rn(s.amb[TM + slot], 1.0f), 2.0f);
                w = __fdiv_rn(__fadd_rn(s.amb[2 * TM + slot], 1.0f), 2.0f);
            }
            {   // parked position features back into tile 0, k[0,32)
                float p16[16];
                tmem_ld16(tmem + (lane_base << 16) + TMEM_P + (uint32_t)(lg * 16), p16);
                wait_ld();
                store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 2 * lg), &p16[0]);
                store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 2 * lg + 1), &p16[8]);
            }
            if (a.amb_octs) {
                float f[16];
                if (v) lookup8o(&s.lvl[1][lg * 16], a.amb_gm.align_off, a.amb_gm.interp == 1, a.amb_octs, u, vv, w, f);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = 0.f;
                }
                store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 4 + 2 * lg), &f[0]);
                store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 4 + 2 * lg + 1), &f[8]);
            } else {
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    float f[8];
                    if (v) lookup4(a.amb_gm, a.amb_tab, a.amb_quads, lg * 8 + c * 4, u, vv, w, f);
                    else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) f[i] = 0.f;
                    }
                    store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 4 + 2 * lg + c), f);
                }
            }
        }
        fence_async_smem(); fence_before_sync(); __syncthreads();
        PH(7)   // ambient gather + unpark

        // ---- sigma net 64 -> 128 -> 128 -> (128 geo + sigma) ----
        if (warp == 0) issue_layer<SPLIT, 4, 1>(s, tmem, idesc128);
        wait_acc<SPLIT, 4, 1>(t, s, acc_uses, warp);
        epilogue_wide<BF16, SPLIT, true, false, false>(&s.a_hi[0][0], &s.a_lo[0][0], tmem, tid, nullptr, nullptr, nullptr);
        fence_async_smem(); fence_before_sync(); __syncthreads();
        if (warp == 0) issue_layer<SPLIT, 5, 2>(s, tmem, idesc128);
        wait_acc<SPLIT, 5, 2>(t, s, acc_uses, warp);
        epilogue_wide<BF16, SPLIT, true, false, false>(&s.a_hi[0][0], &s.a_lo[0][0], tmem, tid, nullptr, nullptr, nullptr);
        fence_async_smem(); fence_before_sync(); __syncthreads();
        if (warp == 0) issue_layer<SPLIT, 7, 2>(s, tmem, idesc144);
        wait_acc<SPLIT, 7, 2>(t, s, acc_uses, warp);
        if (tid < TM) {   // column 128 = sigma logit (weight row 0 was packed last); sigma = density_scale * exp(h)
            float o[16];
            tmem_ld16(tmem + (lane_base << 16) + 128u, o);
            wait_ld();
            s.sig[tid] = a.density_scale * expf(o[0]);
            // SH(dir) -> K16 tile
            float sh[16];
            sh4(sl.g.dx, sl.g.dy, sl.g.dz, sh);
            if (!sl.active) {
#pragma unroll
                for (int i = 0; i < 16; ++i) sh[i] = 0.f;
            }
            store_chunk<BF16, SPLIT>(s.s_hi, s.s_lo, k16_off(tid, 0), &sh[0]);
            store_chunk<BF16, SPLIT>(s.s_hi, s.s_lo, k16_off(tid, 1), &sh[8]);
        }
        epilogue_wide<BF16, SPLIT, false, false, false>(&s.a_hi[0][0], &s.a_lo[0][0], tmem, tid, nullptr, nullptr, nullptr);   // geo features -> tiles 0,1 (k = 0..127)
        fence_async_smem(); fence_before_sync(); __syncthreads();

        PH(8)   // sigma net: 3 MMA layers + 3 epilogues
        // ---- color net (128 geo + 16 SH [+ folded individual code]) -> 128 -> 3 ----
        if (warp == 0) issue_layer<SPLIT, 9, 3>(s, tmem, idesc128);
        wait_acc<SPLIT, 9, 3>(t, s, acc_uses, warp);
        // color L0 epilogue fused with the 3-wide rgb output layer
        epilogue_wide<BF16, SPLIT, true, true, true>(&s.a_hi[0][0], &s.a_lo[0][0], tmem, tid, s.bias, s.nw + 3 * 128, s.part);
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        if (tid < TM) {
            s.rgb[tid] = 1.0f / (1.0f + expf(-(s.part[0 * TM + tid] + s.part[3 * TM + tid])));
            s.rgb[TM + tid] = 1.0f / (1.0f + expf(-(s.part[1 * TM + tid] + s.part[4 * TM + tid])));
            s.rgb[2 * TM + tid] = 1.0f / (1.0f + expf(-(s.part[2 * TM + tid] + s.part[5 * TM + tid])));
        }
        __syncthreads();

        PH(9)   // color net: 2 MMA layers + epilogue + sigmoid
        // ================= composite + advance (head_common.cuh) =================
        composite_and_advance(a, s, sl, mc, tid);
        PH(10)  // composite + march to next sample
        if (a.phase_cycles && tid == 0) atomicAdd(a.phase_cycles + 31, 1ull);   // batches
        // no barrier needed here: the refill starts with one before shared memory is touched again
    }

    // drain the prefetched weight tiles, then release TMEM
    for (uint32_t q = 0; q < W_NSTAGE; ++q) mbar_wait(&s.bar_full[q], 0);   // tiles 0..2 of a batch that never comes
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

static_assert(sizeof(SmemTC<false>) + 1024 <= 113 * 1024, "single-pass kernel must fit twice per SM (228 KB - 2 x 1 KB reserved)");
static_assert(sizeof(SmemTC<true>) + 1024 <= 227 * 1024, "split kernel exceeds the per-CTA shared memory limit");
size_t head_tc_smem_bytes(bool split) { return (split ? sizeof(SmemTC<true>) : sizeof(SmemTC<false>)) + 1024; }

cudaError_t launch_head_tc(const HeadArgs &a, const HeadTcArgs &t, int precision, int total_hint, cudaStream_t st) {
    const bool split = precision == BF16_X3;
    int blocks = sm_count() * (split ? 1 : 2);   // single-pass kernels: two resident CTAs per SM
    if (getenv("GFPP_ONE_CTA")) blocks = sm_count();   // diagnostics: phase timings without the co-resident CTA
    if (total_hint >= 0) {
        const int need = (total_hint + TM - 1) / TM;
        if (need < blocks) blocks = need > 0 ? need : 1;
    }
    const size_t smem = head_tc_smem_bytes(split);
    cudaError_t e;
#define GO(BF, SP)                                                                                             \
    e = cudaFuncSetAttribute(k_head_tc<BF, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
    if (e != cudaSuccess) return e;                                                                            \
    e = cudaFuncSetAttribute(k_head_tc<BF, SP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
    if (e != cudaSuccess) return e;                                                                            \
    if (getenv("GFPP_DEBUG")) {                                                                                \
        int occ = 0;                                                                                           \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_head_tc<BF, SP>, NT, smem);                      \
        fprintf(stderr, "[gfpp] k_head_tc<%d,%d>: %zu B smem, %d CTA/SM, grid %d\n", (int)BF, (int)SP, smem, occ, blocks); \
    }                                                                                                          \
    k_head_tc<BF, SP><<<blocks, NT, smem, st>>>(a, t);
    if (precision == FP16_X1) { GO(false, false) }
    else if (precision == BF16_X1) { GO(true, false) }
    else if (precision == BF16_X3) { GO(true, true) }
    else return cudaErrorInvalidValue;
#undef GO
    return cudaGetLastError();
}

}  // namespace gfpp
