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
#include "common.cuh"
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
constexpr int REFILL_ITERS = 3;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t TMEM_P = 160;     // 32 columns of parked position features

struct SmemTC {
    unsigned char a_hi[2][A_TILE];
    unsigned char a_lo[2][A_TILE];
    unsigned char s_hi[A_K16];
    unsigned char s_lo[A_K16];
    unsigned char w_hi[W_NSTAGE][W_HALF];
    unsigned char w_lo[W_NSTAGE][W_HALF];
    unsigned char n_hi[4][2048];      // resident narrow weights: ambient-out k-tiles 0,1; color-out k-tiles 0,1 ([16 x 64] SW128)
    unsigned char n_lo[4][2048];
    float bias[128];                  // color L0 bias (individual code folded)
    float sx[TM], sy[TM], sz[TM];
    float amb[3 * TM];
    float sig[TM];
    float rgb[3 * TM];
    int frame[TM];
    int valid[TM];
    unsigned long long bar_full[W_NSTAGE];
    unsigned long long bar_acc;
    uint32_t tmem_base;
    int next, end, done;
};

struct Stream {
    uint32_t consumed;   // weight chunks consumed so far (uniform across the CTA)
    uint32_t acc_uses;   // completed accumulator hand-offs
};

template <bool SPLIT>
__device__ __forceinline__ void issue_load(const HeadTcArgs &t, SmemTC &s, uint32_t seq) {
    const uint32_t stage = seq % W_NSTAGE, id = seq % HEAD_TC_NCHUNK;
    const uint32_t bytes = (uint32_t)t.chunk_bytes[id];
    mbar_expect_tx(&s.bar_full[stage], SPLIT ? 2 * bytes : bytes);
    bulk_g2s(s.w_hi[stage], t.w_hi + t.chunk_off[id], bytes, &s.bar_full[stage]);
    if (SPLIT) bulk_g2s(s.w_lo[stage], t.w_lo + t.chunk_off[id], bytes, &s.bar_full[stage]);
}

// Thread 0: issue one layer = `nchunks` streamed weight tiles against A tiles 0,1 (and the K16 SH tile for a k16 chunk).
template <bool SPLIT>
__device__ __forceinline__ void issue_layer(const HeadTcArgs &t, SmemTC &s, Stream &st, int nchunks, uint32_t d_tmem, uint32_t idesc) {
    fence_after_sync();
    uint32_t seq = st.consumed;
    for (int c = 0; c < nchunks; ++c, ++seq) {
        const uint32_t stage = seq % W_NSTAGE, id = seq % HEAD_TC_NCHUNK;
        mbar_wait(&s.bar_full[stage], (seq / W_NSTAGE) & 1);
        fence_after_sync();
        const bool k16 = t.chunk_k16[id] != 0;
        const uint32_t ah = k16 ? smem_u32(s.s_hi) : smem_u32(s.a_hi[c]);
        const uint32_t al = k16 ? smem_u32(s.s_lo) : smem_u32(s.a_lo[c]);
        issue_ktile(d_tmem, ah, al, smem_u32(s.w_hi[stage]), smem_u32(s.w_lo[stage]), t.chunk_ksteps[id], k16, SPLIT, idesc, c > 0);
    }
    mma_commit(&s.bar_acc);
}

// Thread 0: a narrow (N=16) layer from the resident weights, K = 128 (two SW128 k-tiles)
template <bool SPLIT>
__device__ __forceinline__ void issue_narrow(SmemTC &s, int which, uint32_t d_tmem, uint32_t idesc16) {
    fence_after_sync();
    for (int c = 0; c < 2; ++c)
        issue_ktile(d_tmem, smem_u32(s.a_hi[c]), smem_u32(s.a_lo[c]), smem_u32(s.n_hi[which * 2 + c]), smem_u32(s.n_lo[which * 2 + c]), 4,
                    false, SPLIT, idesc16, c > 0);
    mma_commit(&s.bar_acc);
}

// All threads: wait for the accumulator; thread 0 then refills the weight stages the layer has released.
template <bool SPLIT>
__device__ __forceinline__ void wait_acc(const HeadTcArgs &t, SmemTC &s, Stream &st, int nchunks, int tid) {
    mbar_wait(&s.bar_acc, st.acc_uses & 1);
    st.acc_uses += 1;
    fence_after_sync();
    if (tid == 0)
        for (int c = 0; c < nchunks; ++c) issue_load<SPLIT>(t, s, st.consumed + c + W_NSTAGE);
    st.consumed += nchunks;
}

// Epilogue of a 128-wide layer: thread (row, half) reads 64 accumulator columns, applies bias/ReLU and writes them as the
// next layer's A operand (k = column) into tile `half`.
template <bool BF16, bool SPLIT, bool RELU>
__device__ __forceinline__ void epilogue_wide(SmemTC &s, uint32_t tmem, int tid, const float *bias) {
    const int row = tid & 127, half = tid >> 7;
    const uint32_t lane_base = (uint32_t)((tid >> 5) & 3) * 32u;
    const uint32_t taddr = tmem + (lane_base << 16) + (uint32_t)half * 64u;
    float v[4][16];
#pragma unroll
    for (int q = 0; q < 4; ++q) tmem_ld16(taddr + q * 16, v[q]);
    wait_ld();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float x = v[q][i];
            if (bias) x += bias[half * 64 + q * 16 + i];
            if (RELU) x = fmaxf(x, 0.f);
            v[q][i] = x;
        }
        store_chunk<BF16, SPLIT>(s.a_hi[half], s.a_lo[half], sw128_off(row, 2 * q), &v[q][0]);
        store_chunk<BF16, SPLIT>(s.a_hi[half], s.a_lo[half], sw128_off(row, 2 * q + 1), &v[q][8]);
    }
}

}  // namespace

template <bool BF16, bool SPLIT>
__global__ void __launch_bounds__(HEAD_NT, 1) k_head_tc(const __grid_constant__ HeadArgs a, const __grid_constant__ HeadTcArgs t) {
    extern __shared__ __align__(1024) unsigned char smem_raw_[];
    unsigned char *smem_raw = smem_raw_ + ((1024u - (smem_u32(smem_raw_) & 1023u)) & 1023u);
    SmemTC &s = *reinterpret_cast<SmemTC *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- one-time setup ----
    MarchConst mc = a.mc;
    float occ_lo[3] = {0.f, 0.f, 0.f}, occ_hi[3] = {0.f, 0.f, 0.f};
    bool have_box = false;
    if (a.occ_bounds) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { mc.bb_lo[k] = a.occ_bounds[k]; mc.bb_hi[k] = a.occ_bounds[3 + k]; }
        if (a.use_occ_box) {
            have_box = true;
            const float mb = fminf(1.0f, mc.bound);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (mc.bb_hi[k] < mc.bb_lo[k]) { occ_lo[k] = 1e30f; occ_hi[k] = -1e30f; }
                else {
                    occ_lo[k] = ((float)(mc.bb_lo[k] - 1) * mc.rH * 2.0f - 1.0f) * mb;
                    occ_hi[k] = ((float)(mc.bb_hi[k] + 2) * mc.rH * 2.0f - 1.0f) * mb;
                }
            }
        }
    }
    if (warp == 0) tmem_alloc(&s.tmem_base, TMEM_COLS);
    if (tid == 32) {
        for (int i = 0; i < W_NSTAGE; ++i) mbar_init(&s.bar_full[i], 1);
        mbar_init(&s.bar_acc, 1);
        mbar_fence_init();
        s.next = 0; s.end = 0; s.done = 0;
    }
    for (int i = tid; i < 4 * 2048 / 16; i += NT) {
        reinterpret_cast<uint4 *>(&s.n_hi[0][0])[i] = reinterpret_cast<const uint4 *>(t.narrow_hi)[i];
        if (SPLIT) reinterpret_cast<uint4 *>(&s.n_lo[0][0])[i] = reinterpret_cast<const uint4 *>(t.narrow_lo)[i];
    }
    if (tid < 128) s.bias[tid] = a.narrow[7 * 128 + tid];
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s.tmem_base;
    if (tid == 0)
        for (uint32_t q = 0; q < W_NSTAGE; ++q) issue_load<SPLIT>(t, s, q);
    Stream st;
    st.consumed = 0;
    st.acc_uses = 0;
    const uint32_t idesc128 = make_idesc(BF16 ? 1 : 0, 128), idesc144 = make_idesc(BF16 ? 1 : 0, 144), idesc16 = make_idesc(BF16 ? 1 : 0, 16);
    const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;

    const int total = (a.pass == 1) ? a.n_frames * a.n_rays : *a.n_survivors;
    Slot sl;
    sl.active = false;
    sl.gid = 0; sl.frame = 0; sl.nsamp = 0; sl.cap = 0;

    for (;;) {
        // ================= refill dead slots from the global cursor (as in head_kernel.cu) =================
        for (int it = 0; it < REFILL_ITERS; ++it) {
            if (tid == 0 && s.next >= s.end && !s.done) {
                const int base = atomicAdd(a.cursor, TM);
                if (base >= total) { s.done = 1; }
                else { s.next = base; s.end = min(base + TM, total); }
            }
            __syncthreads();
            if (tid < TM && !sl.active && s.next < s.end) {
                const int w = atomicAdd(&s.next, 1);
                if (w < s.end) {
                    int gid = w;
                    if (a.pass == 2) gid = a.survivors[w];
                    sl.gid = gid;
                    sl.frame = gid / a.n_rays;
                    const int ray = gid - sl.frame * a.n_rays;
                    load_ray(a, sl.frame, ray, sl.g);
                    near_far(sl.g, a.aabb, a.min_near, sl.near, sl.far);
                    bool live;
                    if (a.pass == 1) {
                        sl.t = sl.near; sl.ws = 0.f; sl.depth = 0.f; sl.r = sl.gch = sl.b = 0.f;
                        sl.nsamp = 0; sl.cap = a.max_steps;
                        live = may_hit_occupied(have_box, occ_lo, occ_hi, sl.g, sl.near, sl.far) &&
                               march_next(mc, sl.g, sl.far, sl.t, sl.px, sl.py, sl.pz, sl.dt);
                        if (!live) {
                            finalize_ray(a, sl, true);
                            warp_agg_add(a.hist, sl.frame * (a.max_steps + 2) + 1, 1);
                        }
                    } else {
                        const size_t g = (size_t)gid;
                        sl.t = a.rays_t[g]; sl.ws = a.wsum[g]; sl.depth = a.depth[g];
                        sl.r = a.image[3 * g]; sl.gch = a.image[3 * g + 1]; sl.b = a.image[3 * g + 2];
                        sl.nsamp = a.max_steps; sl.cap = a.B_total[sl.frame];
                        live = sl.nsamp < sl.cap && march_next(mc, sl.g, sl.far, sl.t, sl.px, sl.py, sl.pz, sl.dt);
                        if (!live) finalize_ray(a, sl, true);
                    }
                    sl.active = live;
                }
            }
            __syncthreads();
        }

        // ================= publish the batch =================
        if (tid < TM) {
            s.valid[tid] = sl.active ? 1 : 0;
            s.frame[tid] = sl.frame;
            s.sx[tid] = sl.px; s.sy[tid] = sl.py; s.sz[tid] = sl.pz;
        }
        const int n_valid = __syncthreads_count(tid < TM && sl.active);
        if (n_valid == 0) {
            const bool out_of_work = s.done && s.next >= s.end;
            __syncthreads();
            if (out_of_work) break;
            continue;
        }

        const int slot = tid & (TM - 1), lg = tid >> 7;
        const bool v = s.valid[slot] != 0;
        // ---- position grid -> A tile0 k[0,32) and TMEM park; conditioning -> tile0 k[32,64), tile1 k[0,32) ----
        {
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = 0.f;
            float c32[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) c32[i] = 0.f;
            if (v) {
                const float inv2b = 2.0f * mc.bound;
                const float u = __fdiv_rn(__fadd_rn(s.sx[slot], mc.bound), inv2b);
                const float vv = __fdiv_rn(__fadd_rn(s.sy[slot], mc.bound), inv2b);
                const float w = __fdiv_rn(__fadd_rn(s.sz[slot], mc.bound), inv2b);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 g2 = grid_lookup3(a.pos_gm, a.pos_tab, lg * 8 + j, u, vv, w);
                    f[2 * j] = g2.x; f[2 * j + 1] = g2.y;
                }
                const float4 *cf = reinterpret_cast<const float4 *>(a.cond_feat + (size_t)s.frame[slot] * 64 + lg * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 q = __ldg(cf + j);
                    c32[4 * j] = q.x; c32[4 * j + 1] = q.y; c32[4 * j + 2] = q.z; c32[4 * j + 3] = q.w;
                }
            }
            store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 2 * lg), &f[0]);
            store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 2 * lg + 1), &f[8]);
            tmem_st16(tmem + (lane_base << 16) + TMEM_P + (uint32_t)lg * 16u, f);
            // cond values [lg*32, lg*32+32) sit at k = 32 + lg*32 + i: lg 0 -> tile0 chunks 4..7, lg 1 -> tile1 chunks 0..3
#pragma unroll
            for (int j = 0; j < 4; ++j)
                store_chunk<BF16, SPLIT>(s.a_hi[lg], s.a_lo[lg], sw128_off(slot, (lg == 0 ? 4 : 0) + j), &c32[8 * j]);
            wait_st();
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();

        // ---- ambient net 96 -> 128 -> 128 -> 3 ----
        if (tid == 0) issue_layer<SPLIT>(t, s, st, 2, tmem, idesc128);
        wait_acc<SPLIT>(t, s, st, 2, tid);
        epilogue_wide<BF16, SPLIT, true>(s, tmem, tid, nullptr);
        fence_async_smem(); fence_before_sync(); __syncthreads();
        if (tid == 0) issue_layer<SPLIT>(t, s, st, 2, tmem, idesc128);
        wait_acc<SPLIT>(t, s, st, 2, tid);
        epilogue_wide<BF16, SPLIT, true>(s, tmem, tid, nullptr);
        fence_async_smem(); fence_before_sync(); __syncthreads();
        if (tid == 0) issue_narrow<SPLIT>(s, 0, tmem, idesc16);
        wait_acc<SPLIT>(t, s, st, 0, tid);
        if (tid < TM) {
            float o[16];
            tmem_ld16(tmem + (lane_base << 16), o);
            wait_ld();
            s.amb[tid] = tanhf(o[0]);
            s.amb[TM + tid] = tanhf(o[1]);
            s.amb[2 * TM + tid] = tanhf(o[2]);
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        // ---- sigma-net input: tile0 k[0,32) <- parked position features, k[32,64) <- ambient grid ----
        {
            float f[16];
            tmem_ld16(tmem + (lane_base << 16) + TMEM_P + (uint32_t)lg * 16u, f);
            wait_ld();
            store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 2 * lg), &f[0]);
            store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 2 * lg + 1), &f[8]);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = 0.f;
            if (v) {
                const float u = __fdiv_rn(__fadd_rn(s.amb[slot], 1.0f), 2.0f);
                const float vv = __fdiv_rn(__fadd_rn(s.amb[TM + slot], 1.0f), 2.0f);
                const float w = __fdiv_rn(__fadd_rn(s.amb[2 * TM + slot], 1.0f), 2.0f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 g2 = (a.amb_gm.dim == 3) ? grid_lookup3(a.amb_gm, a.amb_tab, lg * 8 + j, u, vv, w)
                                                          : grid_lookup2(a.amb_gm, a.amb_tab, lg * 8 + j, u, vv);
                    f[2 * j] = g2.x; f[2 * j + 1] = g2.y;
                }
            }
            store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 4 + 2 * lg), &f[0]);
            store_chunk<BF16, SPLIT>(s.a_hi[0], s.a_lo[0], sw128_off(slot, 4 + 2 * lg + 1), &f[8]);
        }
        fence_async_smem(); fence_before_sync(); __syncthreads();

        // ---- sigma net 64 -> 128 -> 128 -> (128 geo + sigma) ----
        if (tid == 0) issue_layer<SPLIT>(t, s, st, 1, tmem, idesc128);
        wait_acc<SPLIT>(t, s, st, 1, tid);
        epilogue_wide<BF16, SPLIT, true>(s, tmem, tid, nullptr);
        fence_async_smem(); fence_before_sync(); __syncthreads();
        if (tid == 0) issue_layer<SPLIT>(t, s, st, 2, tmem, idesc128);
        wait_acc<SPLIT>(t, s, st, 2, tid);
        epilogue_wide<BF16, SPLIT, true>(s, tmem, tid, nullptr);
        fence_async_smem(); fence_before_sync(); __syncthreads();
        if (tid == 0) issue_layer<SPLIT>(t, s, st, 2, tmem, idesc144);
        wait_acc<SPLIT>(t, s, st, 2, tid);
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
        epilogue_wide<BF16, SPLIT, false>(s, tmem, tid, nullptr);   // geo features -> tiles 0,1 (k = 0..127)
        fence_async_smem(); fence_before_sync(); __syncthreads();

        // ---- color net (128 geo + 16 SH [+ folded individual code]) -> 128 -> 3 ----
        if (tid == 0) issue_layer<SPLIT>(t, s, st, 3, tmem, idesc128);
        wait_acc<SPLIT>(t, s, st, 3, tid);
        epilogue_wide<BF16, SPLIT, true>(s, tmem, tid, s.bias);
        fence_async_smem(); fence_before_sync(); __syncthreads();
        if (tid == 0) issue_narrow<SPLIT>(s, 1, tmem, idesc16);
        wait_acc<SPLIT>(t, s, st, 0, tid);
        if (tid < TM) {
            float o[16];
            tmem_ld16(tmem + (lane_base << 16), o);
            wait_ld();
            s.rgb[tid] = 1.0f / (1.0f + expf(-o[0]));
            s.rgb[TM + tid] = 1.0f / (1.0f + expf(-o[1]));
            s.rgb[2 * TM + tid] = 1.0f / (1.0f + expf(-o[2]));
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();

        // ================= composite + advance (identical to head_kernel.cu) =================
        if (tid < TM && sl.active) {
            const float sigma = s.sig[tid];
            const float alpha = 1.0f - expf(-sigma * sl.dt);
            const float T = 1.0f - sl.ws;
            const float w = alpha * T;
            sl.ws += w;
            sl.depth += w * sl.t;
            sl.r += w * s.rgb[tid];
            sl.gch += w * s.rgb[TM + tid];
            sl.b += w * s.rgb[2 * TM + tid];
            sl.nsamp += 1;
            if (a.valid_samples) warp_agg_add(a.valid_samples, sl.frame, 1);
            int D = 0;
            bool suspend = false;
            if (T < a.T_thresh) D = sl.nsamp;
            else if (sl.nsamp >= sl.cap) suspend = true;
            else if (!march_next(mc, sl.g, sl.far, sl.t, sl.px, sl.py, sl.pz, sl.dt)) D = sl.nsamp + 1;
            if (D) {
                finalize_ray(a, sl, true);
                if (a.pass == 1) warp_agg_add(a.hist, sl.frame * (a.max_steps + 2) + D, 1);
                sl.active = false;
            } else if (suspend) {
                if (a.pass == 1) {
                    finalize_ray(a, sl, false);
                    a.rays_t[sl.gid] = sl.t;
                    cg::coalesced_group grp = cg::coalesced_threads();
                    int base = 0;
                    if (grp.thread_rank() == 0) base = atomicAdd(a.n_survivors, (int)grp.size());
                    base = grp.shfl(base, 0);
                    a.survivors[base + grp.thread_rank()] = sl.gid;
                } else {
                    finalize_ray(a, sl, true);
                }
                sl.active = false;
            }
        }
    }

    // drain the prefetched weight tiles, then release TMEM
    for (uint32_t q = 0; q < W_NSTAGE; ++q) {
        const uint32_t seq = st.consumed + q;
        mbar_wait(&s.bar_full[seq % W_NSTAGE], (seq / W_NSTAGE) & 1);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

size_t head_tc_smem_bytes() { return sizeof(SmemTC) + 1024; }

cudaError_t launch_head_tc(const HeadArgs &a, const HeadTcArgs &t, int precision, int total_hint, cudaStream_t st) {
    int blocks = sm_count();
    if (total_hint >= 0) {
        const int need = (total_hint + TM - 1) / TM;
        if (need < blocks) blocks = need > 0 ? need : 1;
    }
    const size_t smem = head_tc_smem_bytes();
    cudaError_t e;
#define GO(BF, SP)                                                                                             \
    e = cudaFuncSetAttribute(k_head_tc<BF, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
    if (e != cudaSuccess) return e;                                                                            \
    k_head_tc<BF, SP><<<blocks, NT, smem, st>>>(a, t);
    if (precision == FP16_X1) { GO(false, false) }
    else if (precision == BF16_X1) { GO(true, false) }
    else if (precision == BF16_X3) { GO(true, true) }
    else return cudaErrorInvalidValue;
#undef GO
    return cudaGetLastError();
}

}  // namespace gfpp
