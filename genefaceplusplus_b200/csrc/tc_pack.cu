// tc_pack.cu -- weight repacking for the tcgen05 path + a one-tile GEMM self-test of the tcgen05 plumbing.
#include "launch.cuh"
#include "tc.cuh"

namespace gfpp {

using namespace tc;

// One K-tile of nn.Linear weights W[.][ld] (fp32) -> 16-bit hi (and lo) tiles in the UMMA K-major layout, ready to be
// bulk-copied into shared memory: tile row dst_row0+n, column k  <-  W[row0+n][col0+k] for n < N, k < kc.
// The destination must be zero-initialised (padding rows / columns stay zero).
__global__ void k_pack_tc_tile(const float *__restrict__ W, int ld, int row0, int col0, int N, int kc, int dst_row0,
                               int k16, int bf16, unsigned char *__restrict__ hi, unsigned char *__restrict__ lo) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * kc; i += gridDim.x * blockDim.x) {
        const int n0 = i / kc, k = i - n0 * kc, n = dst_row0 + n0;
        const float v = W[(size_t)(row0 + n0) * ld + col0 + k];
        const uint32_t off = (k16 ? k16_off(n, k >> 3) : sw128_off(n, k >> 3)) + (uint32_t)(k & 7) * 2u;
        if (bf16) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            *reinterpret_cast<__nv_bfloat16 *>(hi + off) = h;
            if (lo) *reinterpret_cast<__nv_bfloat16 *>(lo + off) = __float2bfloat16_rn(v - __bfloat162float(h));
        } else {
            const __half h = __float2half_rn(v);
            *reinterpret_cast<__half *>(hi + off) = h;
            if (lo) *reinterpret_cast<__half *>(lo + off) = __float2half_rn(v - __half2float(h));
        }
    }
}

cudaError_t launch_pack_tc_tile(const float *W, int ld, int row0, int col0, int N, int kc, int dst_row0, int k16, int bf16,
                                unsigned char *hi, unsigned char *lo, cudaStream_t st) {
    const int total = N * kc;
    k_pack_tc_tile<<<grid_for((uint64_t)total, 256), 256, 0, st>>>(W, ld, row0, col0, N, kc, dst_row0, k16, bf16, hi, lo);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Self-test: out[128][N] = A[128][K] * W[N][K]^T through exactly the tile layouts, descriptors, MMA issue, commit and
// TMEM read-back the fused kernel uses.  K = 64*nt64 (+16 when k16_tail); the weight tiles come pre-packed.
struct TcTestSmem {
    unsigned char a_hi[3][16384];
    unsigned char a_lo[3][16384];
    unsigned char w_hi[3][18432];
    unsigned char w_lo[3][18432];
    unsigned long long bar_w, bar_acc;
    uint32_t tmem_base;
};

template <bool BF16, bool SPLIT>
__global__ void __launch_bounds__(256, 1) k_tc_selftest(const float *__restrict__ A, int K, const unsigned char *__restrict__ wt_hi,
                                                        const unsigned char *__restrict__ wt_lo, int N, int nt64, int k16_tail,
                                                        float *__restrict__ out) {
    extern __shared__ __align__(1024) unsigned char raw_[];
    unsigned char *raw = raw_ + ((1024u - (smem_u32(raw_) & 1023u)) & 1023u);  // SW128 tiles need 1024-byte alignment
    TcTestSmem &s = *reinterpret_cast<TcTestSmem *>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int ntiles = nt64 + (k16_tail ? 1 : 0);
    if (warp == 0) tmem_alloc(&s.tmem_base, 256);
    if (tid == 32) {
        mbar_init(&s.bar_w, 1);
        mbar_init(&s.bar_acc, 1);
        mbar_fence_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s.tmem_base;
    const uint32_t tile_bytes = 18432;
    if (tid == 0) {
        uint32_t bytes = 0;
        for (int t = 0; t < ntiles; ++t) bytes += tile_bytes * (SPLIT ? 2 : 1);
        mbar_expect_tx(&s.bar_w, bytes);
        for (int t = 0; t < ntiles; ++t) {
            bulk_g2s(s.w_hi[t], wt_hi + (size_t)t * tile_bytes, tile_bytes, &s.bar_w);
            if (SPLIT) bulk_g2s(s.w_lo[t], wt_lo + (size_t)t * tile_bytes, tile_bytes, &s.bar_w);
        }
    }
    // A: fp32 global -> 16-bit swizzled tiles; thread (row, part) converts chunks part, part+2, ...
    {
        const int row = tid & 127, part = tid >> 7;
        for (int c = part; c < K / 8; c += 2) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = A[(size_t)row * K + c * 8 + i];
            const int t = c / 8, cc = c % 8;
            const bool is16 = k16_tail && t == nt64;
            const uint32_t off = is16 ? k16_off(row, cc) : sw128_off(row, cc);
            store_chunk<BF16, SPLIT>(s.a_hi[t], s.a_lo[t], off, v);
        }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        mbar_wait(&s.bar_w, 0);
        fence_after_sync();
        const uint32_t idesc = make_idesc(BF16 ? 1 : 0, N);
        for (int t = 0; t < ntiles; ++t) {
            const bool is16 = k16_tail && t == nt64;
            // a partial last 64-wide tile only holds (K - 64*t)/16 k-steps of defined operand data
            const int k64 = K - (k16_tail ? 16 : 0);
            const int ks = is16 ? 1 : ((k64 - 64 * t) >= 64 ? 4 : (k64 - 64 * t) / 16);
            issue_ktile(tmem, smem_u32(s.a_hi[t]), smem_u32(s.a_lo[t]), smem_u32(s.w_hi[t]), smem_u32(s.w_lo[t]), ks, is16,
                        SPLIT, idesc, t > 0);
        }
        mma_commit(&s.bar_acc);
    }
    mbar_wait(&s.bar_acc, 0);
    fence_after_sync();
    if (tid < 128) {
        const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
        for (int c = 0; c < N / 16; ++c) {
            float v[16];
            tmem_ld16(tmem + (lane_base << 16) + (uint32_t)c * 16u, v);
            wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) out[(size_t)tid * N + c * 16 + i] = v[i];
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

cudaError_t launch_tc_selftest(const float *A, int K, const unsigned char *wt_hi, const unsigned char *wt_lo, int N, int nt64,
                               int k16_tail, int precision, float *out, cudaStream_t st) {
    const size_t smem = sizeof(TcTestSmem) + 1024;
    cudaError_t e;
#define GO(BF, SP)                                                                                                     \
    e = cudaFuncSetAttribute(k_tc_selftest<BF, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    if (e != cudaSuccess) return e;                                                                                    \
    k_tc_selftest<BF, SP><<<1, 256, smem, st>>>(A, K, wt_hi, wt_lo, N, nt64, k16_tail, out);
    if (precision == FP16_X1) { GO(false, false) }
    else if (precision == BF16_X1) { GO(true, false) }
    else if (precision == BF16_X3) { GO(true, true) }
    else return cudaErrorInvalidValue;
#undef GO
    return cudaGetLastError();
}

}  // namespace gfpp
