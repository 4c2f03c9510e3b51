// torso_kernel.cu -- torso field + final composite (the epilogue of every frame).
//
// Replaces radnerf_torso.py:156-197 (mask by grid_sample of density_grid_torso, forward_torso on the
// masked pixels :51-84, three-way composite, clamp) and, for the head-only model, renderer.py:386-392.
// One persistent kernel: a CTA takes tiles of 128 consecutive pixels of one frame; tiles with no torso
// pixel skip the MLPs.  Per-frame constants (freq-encoded pose, torso code) are folded into per-frame
// bias vectors by k_torso_frame_bias, so the per-pixel GEMMs only see the pixel-dependent columns
// (deform: 42 of 104, canonical: 74 of 136).
#include "common.cuh"
#include "launch.cuh"
#include "torso_common.cuh"
#include "torso_kernel.cuh"

namespace gfpp {

using namespace torsoc;

namespace {

constexpr int TP = 128;    // pixels per tile
constexpr int NT = 256;
constexpr int LDT = 84;    // A-tile stride (>= 76, 16B aligned, 84 % 32 = 20)
constexpr int LDE = 44;    // enc_x copy stride
constexpr int KD0 = 44, KC0 = 76;

struct Smem {
    float A[TP * LDT];
    float E[TP * LDE];
    float wd0[KD0 * 64], wd1[64 * 64], wc0[KC0 * 32], wc1[32 * 32];
    float wd2[2 * 64], wc2[4 * 32];
    float x2[2 * TP];
    float alpha[TP], col[3 * TP], dxy[2 * TP];
    int mask[TP];
};

}  // namespace

// per-frame bias vectors: columns of the first deform / canonical layers that do not depend on the pixel
__global__ void k_torso_frame_bias(TorsoArgs a, const float *__restrict__ w_def0 /*[64,104]*/,
                                   const float *__restrict__ w_can0 /*[32,136]*/, const float *__restrict__ code,
                                   int code_dim, float *__restrict__ bias_def, float *__restrict__ bias_can) {
    const int f = blockIdx.x, j = threadIdx.x;
    __shared__ float h[64];
    const int n_pose = 6 + 6 * 2 * 4;  // 54
    if (j < n_pose) h[j] = freq_entry(a.pose6 + (size_t)f * 6, 6, j);
    else if (j < n_pose + code_dim) h[j] = code ? code[j - n_pose] : 0.f;
    __syncthreads();
    const int nh = n_pose + code_dim, din = 42 + nh;
    if (j < 64) {
        float s = 0.f;
        for (int k = 0; k < nh; ++k) s = fmaf(w_def0[j * din + 42 + k], h[k], s);
        bias_def[(size_t)f * 64 + j] = s;
    }
    if (j < 32) {
        float s = 0.f;
        for (int k = 0; k < nh; ++k) s = fmaf(w_can0[j * (32 + din) + 32 + 42 + k], h[k], s);
        bias_can[(size_t)f * 32 + j] = s;
    }
}

__global__ void __launch_bounds__(256, 2) k_epilogue(const __grid_constant__ TorsoArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &s = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x;
    const int tiles_per_frame = (a.n_rays + TP - 1) / TP;
    const int n_tiles = a.n_frames * tiles_per_frame;
    if (a.has_torso) {
        for (int i = tid; i < KD0 * 64; i += NT) s.wd0[i] = a.w_def0[i];
        for (int i = tid; i < 64 * 64; i += NT) s.wd1[i] = a.w_def1[i];
        for (int i = tid; i < KC0 * 32; i += NT) s.wc0[i] = a.w_can0[i];
        for (int i = tid; i < 32 * 32; i += NT) s.wc1[i] = a.w_can1[i];
        for (int i = tid; i < 2 * 64; i += NT) s.wd2[i] = a.w_def2[i];
        for (int i = tid; i < 4 * 32; i += NT) s.wc2[i] = a.w_can2[i];
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    const int slot = tid & (TP - 1), lg = tid >> 7;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int f = tile / tiles_per_frame;
        const int n0 = (tile - f * tiles_per_frame) * TP;
        const int n = n0 + slot;
        const bool in_range = n < a.n_rays;
        int m = 0;
        if (a.has_torso && tid < TP && in_range) {
            const float occ = sample_density(a.density_grid_torso, a.grid_size, a.bg_coords[2 * n], a.bg_coords[2 * n + 1]);
            m = occ > a.density_thresh_torso ? 1 : 0;
        }
        if (tid < TP) s.mask[tid] = m;
        const int any = __syncthreads_or(m);
        if (any) {
            // ---- enc_x = freq10(shrink * bg_coord) -> E[:, 0:42] and A[:, 0:42]; pad cols 42,43 = 0 ----
            {
                float x[2] = {0.f, 0.f};
                if (in_range) { x[0] = __fmul_rn(a.bg_coords[2 * n], a.torso_shrink); x[1] = __fmul_rn(a.bg_coords[2 * n + 1], a.torso_shrink); }
                for (int c = lg * 22; c < lg * 22 + 22; ++c) {
                    const float v = c < 42 ? freq_entry(x, 2, c) : 0.f;
                    s.E[slot * LDE + c] = v;
                    s.A[slot * LDT + c] = v;
                }
                if (lg == 0) { s.x2[slot] = x[0]; s.x2[TP + slot] = x[1]; }
            }
            __syncthreads();
            // ---- deform net 104 -> 64 -> 64 -> 2 (radnerf_torso.py:60-76) ----
            {
                float acc[8][4] = {};
                small_gemm<4, KD0, LDT>(acc, s.A, s.wd0, ty, tx);
                __syncthreads();
                store_relu<4, LDT>(acc, s.A, a.bias_def + (size_t)f * 64, ty, tx);
                __syncthreads();
            }
            {
                float acc[8][4] = {};
                small_gemm<4, 64, LDT>(acc, s.A, s.wd1, ty, tx);
                __syncthreads();
                store_relu<4, LDT>(acc, s.A, nullptr, ty, tx);
                __syncthreads();
            }
            {
                const int row = tid >> 1, half = tid & 1;
                float v0 = 0.f, v1 = 0.f;
                for (int k = half * 32; k < half * 32 + 32; ++k) {
                    const float xk = s.A[row * LDT + k];
                    v0 = fmaf(xk, s.wd2[k], v0);
                    v1 = fmaf(xk, s.wd2[64 + k], v1);
                }
                v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
                v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
                if (half == 0) {
                    s.dxy[row] = v0; s.dxy[TP + row] = v1;
                    // x = (x + dx).clamp(-1, 1)
                    s.x2[row] = fminf(fmaxf(s.x2[row] + v0, -1.f), 1.f);
                    s.x2[TP + row] = fminf(fmaxf(s.x2[TP + row] + v1, -1.f), 1.f);
                }
            }
            __syncthreads();
            // ---- canonical input: [tiled2D(x) 32 | enc_x 42 | pad 2] ----
            {
                const float u = __fdiv_rn(__fadd_rn(s.x2[slot], 1.0f), 2.0f), v = __fdiv_rn(__fadd_rn(s.x2[TP + slot], 1.0f), 2.0f);
#pragma unroll 2
                for (int l = lg * 8; l < lg * 8 + 8; ++l)
                    *reinterpret_cast<float2 *>(s.A + slot * LDT + 2 * l) = grid_lookup2(a.tor_gm, a.tor_tab, l, u, v);
                for (int c = lg * 22; c < lg * 22 + 22; ++c) s.A[slot * LDT + 32 + c] = s.E[slot * LDE + c];
            }
            __syncthreads();
            // ---- canonical net 136 -> 32 -> 32 -> 4, sigmoid (radnerf_torso.py:77-82) ----
            {
                float acc[8][2] = {};
                small_gemm<2, KC0, LDT>(acc, s.A, s.wc0, ty, tx);
                __syncthreads();
                store_relu<2, LDT>(acc, s.A, a.bias_can + (size_t)f * 32, ty, tx);
                __syncthreads();
            }
            {
                float acc[8][2] = {};
                small_gemm<2, 32, LDT>(acc, s.A, s.wc1, ty, tx);
                __syncthreads();
                store_relu<2, LDT>(acc, s.A, nullptr, ty, tx);
                __syncthreads();
            }
            {
                const int row = tid >> 1, half = tid & 1;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k = half * 16; k < half * 16 + 16; ++k) {
                    const float xk = s.A[row * LDT + k];
#pragma unroll
                    for (int o = 0; o < 4; ++o) v[o] = fmaf(xk, s.wc2[o * 32 + k], v[o]);
                }
#pragma unroll
                for (int o = 0; o < 4; ++o) v[o] += __shfl_xor_sync(0xffffffffu, v[o], 1);
                if (half == 0) {
                    s.alpha[row] = 1.0f / (1.0f + expf(-v[0]));
                    s.col[row] = 1.0f / (1.0f + expf(-v[1]));
                    s.col[TP + row] = 1.0f / (1.0f + expf(-v[2]));
                    s.col[2 * TP + row] = 1.0f / (1.0f + expf(-v[3]));
                }
            }
            __syncthreads();
        }
        // ---- composite (radnerf_torso.py:187-193 / renderer.py:386-392) ----
        if (tid < TP && in_range) {
            const size_t g = (size_t)f * a.n_rays + n;
            float ta = 0.f, tc[3] = {0.f, 0.f, 0.f}, dx = 0.f, dy = 0.f;
            if (any && s.mask[tid]) {
                ta = s.alpha[tid];
                tc[0] = s.col[tid]; tc[1] = s.col[TP + tid]; tc[2] = s.col[2 * TP + tid];
                dx = s.dxy[tid]; dy = s.dxy[TP + tid];
            }
            const float ws = a.wsum[g];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float bgc = a.bg_color ? a.bg_color[3 * (size_t)n + c] : 1.0f;
                // bg = torso_color * alpha + bg * (1 - alpha); image = image + (1 - ws) * bg; clamp(0, 1)
                const float bg = a.has_torso ? __fadd_rn(__fmul_rn(tc[c], ta), __fmul_rn(bgc, __fsub_rn(1.0f, ta))) : bgc;
                const float v = __fadd_rn(a.image[3 * g + c], __fmul_rn(__fsub_rn(1.0f, ws), bg));
                const float vc = fminf(fmaxf(v, 0.f), 1.f);
                if (a.rgb_map) a.rgb_map[3 * g + c] = vc;
                // (rgb * 255).int() -> uint8: what the driver hands the video writer (genefacepp_infer.py:469, 505)
                if (a.rgb_u8) a.rgb_u8[3 * g + c] = (unsigned char)(int)__fmul_rn(vc, 255.0f);
                if (a.torso_rgb) a.torso_rgb[3 * g + c] = bg;
            }
            if (a.torso_alpha) a.torso_alpha[g] = ta;
            if (a.deform) { a.deform[2 * g] = dx; a.deform[2 * g + 1] = dy; }
            if (a.P_count && any && s.mask[tid]) atomicAdd(a.P_count + f, 1);
        }
        __syncthreads();
    }
}

// dst[k][n] = src[(row0 + n) * ld + col0 + k] for k < K, zero for K <= k < Kpad   (k-major repack of nn.Linear weights)
__global__ void k_pack_kmajor(const float *__restrict__ src, int ld, int row0, int col0, int N, int K, int Kpad,
                              float *__restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Kpad * N; i += gridDim.x * blockDim.x) {
        const int k = i / N, n = i - k * N;
        dst[i] = k < K ? src[(size_t)(row0 + n) * ld + col0 + k] : 0.f;
    }
}

// dst[o][k] = src[(row0 + o) * ld + k]   (plain row copy for the narrow output layers)
__global__ void k_pack_rows(const float *__restrict__ src, int ld, int row0, int n_rows, int K, int dst_ld,
                            float *__restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows * K; i += gridDim.x * blockDim.x) {
        const int o = i / K, k = i - o * K;
        dst[o * dst_ld + k] = src[(size_t)(row0 + o) * ld + k];
    }
}

// bias[n] = sum_k W[n][col0 + k] * v[k]   (folds a per-model constant input block into a bias)
__global__ void k_fold_bias(const float *__restrict__ W, int ld, int col0, int nk, const float *__restrict__ v, int N,
                            float *__restrict__ bias) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int k = 0; k < nk; ++k) s = fmaf(W[(size_t)n * ld + col0 + k], v ? v[k] : 0.f, s);
    bias[n] = s;
}

size_t epilogue_smem_bytes() { return sizeof(Smem); }

cudaError_t launch_torso_frame_bias(const TorsoArgs &a, const float *w_def0, const float *w_can0, const float *code,
                                    int code_dim, float *bias_def, float *bias_can, cudaStream_t st) {
    k_torso_frame_bias<<<a.n_frames, 64, 0, st>>>(a, w_def0, w_can0, code, code_dim, bias_def, bias_can);
    return cudaGetLastError();
}

cudaError_t launch_epilogue(const TorsoArgs &a, cudaStream_t st) {
    // function attributes are per device: set on every launch (a few microseconds), never cached process-wide
    cudaError_t e = cudaFuncSetAttribute(k_epilogue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return e;
    const int tiles = a.n_frames * ((a.n_rays + TP - 1) / TP);
    int blocks = 2 * sm_count();
    if (tiles < blocks) blocks = tiles > 0 ? tiles : 1;
    k_epilogue<<<blocks, NT, sizeof(Smem), st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_pack_kmajor(const float *src, int ld, int row0, int col0, int N, int K, int Kpad, float *dst, cudaStream_t st) {
    k_pack_kmajor<<<grid_for((uint64_t)Kpad * N, 256), 256, 0, st>>>(src, ld, row0, col0, N, K, Kpad, dst);
    return cudaGetLastError();
}
cudaError_t launch_pack_rows(const float *src, int ld, int row0, int n_rows, int K, int dst_ld, float *dst, cudaStream_t st) {
    k_pack_rows<<<grid_for((uint64_t)n_rows * K, 256), 256, 0, st>>>(src, ld, row0, n_rows, K, dst_ld, dst);
    return cudaGetLastError();
}
cudaError_t launch_fold_bias(const float *W, int ld, int col0, int nk, const float *v, int N, float *bias, cudaStream_t st) {
    k_fold_bias<<<(N + 127) / 128, 128, 0, st>>>(W, ld, col0, nk, v, N, bias);
    return cudaGetLastError();
}

}  // namespace gfpp
