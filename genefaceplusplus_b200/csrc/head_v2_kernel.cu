// head_v2_kernel.cu -- fused persistent head renderer, second generation (tcgen05 + TMEM + bulk-copy weight stream).
//
// Replaces the reference's host-driven round loop (modules/radnerfs/renderer.py:340-384: march_rays -> GridEncoder ->
// ambient MLP -> GridEncoder -> sigma MLP -> SHEncoder -> color MLP -> composite_rays, ~40 launches per round) like
// head_tc_kernel.cu does, but organised around ROW OWNERSHIP instead of CTA-wide phases:
//
//   * one CTA per SM; NSLOT "slots" of 128 samples, each served by 4 warps whose thread r owns row r END TO END: it owns
//     the ray (marching, compositing, refill), gathers the row's 2 x 16 grid levels, runs every epilogue of the row
//     (TMEM lane r -> ReLU -> 16-bit operand row of the next layer) and keeps the row's scalars (sample position, ambient
//     coordinates, sigma, rgb) in REGISTERS.  No thread ever reads another thread's row, so the steady state has no
//     __syncthreads and no worker<->worker barrier at all;
//   * one issue warp walks the eight MMA groups of the MLPs layer by layer and, inside a layer, slot by slot: it waits on
//     the slot's `a_ready` mbarrier (4 warp arrivals), issues that slot's tcgen05.mma's against the staged weight tile and
//     commits to the slot's `d_ready` mbarrier.  While slot s's MMAs execute, the owners of the other slots run their
//     epilogues / gathers / ray work: the slots pipeline against each other by construction;
//   * the weight stream is SHARED: a tile is copied once (cp.async.bulk by the loader warp, NSTAGE-deep ring) and used by all
//     NSLOT slots before its stage is released (tcgen05.commit -> `w_free`), which cuts the L2->SM weight traffic per sample
//     by NSLOT -- at this kernel's rate an unshared stream (200 KB per 128 samples) would alone eat half of the ~42 B/clk an
//     SM can pull from L2;
//   * the 3-wide output layers (ambient coordinates, rgb) and the sigma logit are N=16 MMAs against small RESIDENT tiles;
//     the conditioning columns of ambient L0 are folded into a per-frame fp32 bias (k_amb_frame_bias) added in the epilogue.
//
// Precision modes: FP16_X1 (the reference's autocast arithmetic) with 3 slots, and FP16_ROBUST with 2 slots: fp16 hi/lo
// split (3 MMAs per k-step, ~22 mantissa bits) on the ambient net and a 16-bit fixed-point position table (gather8.cuh) --
// the two roundings the field amplifies (tools/error_budget.py) -- everything else as FP16_X1.
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

constexpr int TM = HEAD_TM;          // rows per slot
constexpr uint32_t D_COLS = 144;     // TMEM columns per slot: 128 wide outputs + 16 for the N=16 groups

template <bool ROBUST>
struct Cfg {
    static constexpr int NSLOT = ROBUST ? 2 : 3;
    static constexpr int NSTAGE = 4;   // robust: 4 stages fit with 240 bytes to spare; with 3 the lo-image tiles of ambient L1 arrived late (7.4 K cycles of wait)
    static constexpr int NTILE = ROBUST ? V2_NTILE_ROBUST : V2_NTILE_X1;
    static constexpr int NWORK = 4 * NSLOT;                 // row-owner warps
    static constexpr int NT = 32 * (NWORK + 2);             // + issue warp + loader warp
    static constexpr int SLOT_BYTES = (ROBUST ? 4 : 2) * V2_TILE_BYTES + 4096;   // h0 h1 [l0 l1] sh
    static constexpr int OFF_W = NSLOT * SLOT_BYTES;
    static constexpr int OFF_RES = OFF_W + NSTAGE * V2_TILE_BYTES;
    static constexpr int OFF_MISC = OFF_RES + V2_RES_BYTES;
};

struct Misc {
    float cbias[128];                    // color L0 bias (individual code folded)
    uint32_t coarse[HEAD_COARSE_WORDS];
    uint4 lvl[2][2 * GFPP_MAX_LEVELS];
    unsigned long long a_ready[3], d_ready[3], w_full[4], w_free[4], res_full, batch_bar;
    uint32_t tmem_base;
    volatile int stop;
    volatile int flag[2][16];            // per batch parity, per owner warp: "I have live rows"
};

template <bool ROBUST>
constexpr size_t smem_bytes() { return (size_t)Cfg<ROBUST>::OFF_MISC + sizeof(Misc) + 1024; }
static_assert(smem_bytes<false>() <= 227 * 1024 && smem_bytes<true>() <= 227 * 1024, "v2 shared memory budget");

// ---------------------------------------------------------------- issue-side helpers (warp-uniform operands)
__device__ __forceinline__ void mma_sw(uint32_t d, uint32_t a, uint32_t w, int ksteps, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (k < ksteps) mma_f16(d, desc_sw128(a + (uint32_t)k * 32u), desc_sw128(w + (uint32_t)k * 32u), idesc, (k > 0) ? 1u : acc_first);
}

// ---------------------------------------------------------------- owner-side helpers
__device__ __forceinline__ uint32_t relu_h2(uint32_t h) {
    const __half2 z = __float2half2_rn(0.f);
    const __half2 r = __hmax2(*reinterpret_cast<const __half2 *>(&h), z);
    return *reinterpret_cast<const uint32_t *>(&r);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t h) { return __half22float2(*reinterpret_cast<const __half2 *>(&h)); }

// eight consecutive K values of row `row` -> chunk `chunk` of an SW128 tile (hi) and, when LO, their residuals into `lo`
template <bool RELU, bool LO>
__device__ __forceinline__ void put8(unsigned char *hi, unsigned char *lo, int row, int chunk, const float *v) {
    const uint32_t off = sw128_off(row, chunk);
    uint4 h;
    if (LO) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = RELU ? fmaxf(v[i], 0.f) : v[i];
        h.x = pack2<false>(x[0], x[1]); h.y = pack2<false>(x[2], x[3]); h.z = pack2<false>(x[4], x[5]); h.w = pack2<false>(x[6], x[7]);
        const float2 a = unpack_h2(h.x), b = unpack_h2(h.y), c = unpack_h2(h.z), d = unpack_h2(h.w);
        uint4 l;
        l.x = pack2<false>(x[0] - a.x, x[1] - a.y); l.y = pack2<false>(x[2] - b.x, x[3] - b.y);
        l.z = pack2<false>(x[4] - c.x, x[5] - c.y); l.w = pack2<false>(x[6] - d.x, x[7] - d.y);
        *reinterpret_cast<uint4 *>(lo + off) = l;
    } else {
        h.x = pack2<false>(v[0], v[1]); h.y = pack2<false>(v[2], v[3]); h.z = pack2<false>(v[4], v[5]); h.w = pack2<false>(v[6], v[7]);
        if (RELU) { h.x = relu_h2(h.x); h.y = relu_h2(h.y); h.z = relu_h2(h.z); h.w = relu_h2(h.w); }   // max(rn(x), 0) == rn(max(x, 0))
    }
    *reinterpret_cast<uint4 *>(hi + off) = h;
}

// Epilogue of a 128-wide layer for the calling thread's row: TMEM lane -> (+bias) -> (ReLU) -> operand tiles h0|h1 (cols 0-63 | 64-127).
// BIAS: 0 none, 1 per-frame global vector (ambient L0: the folded conditioning columns), 2 shared-memory vector (color L0).
template <bool RELU, int BIAS, bool LO>
__device__ __forceinline__ void epilogue_row(uint32_t taddr, unsigned char *h, unsigned char *l, int row, const float *bias) {
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)p * 32u, v);
        float4 bq[8];
        if (BIAS) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                bq[j] = BIAS == 1 ? __ldg(reinterpret_cast<const float4 *>(bias) + p * 8 + j) : reinterpret_cast<const float4 *>(bias)[p * 8 + j];
        }
        wait_ld();
        if (BIAS) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { v[4 * j] += bq[j].x; v[4 * j + 1] += bq[j].y; v[4 * j + 2] += bq[j].z; v[4 * j + 3] += bq[j].w; }
        }
        unsigned char *ht = h + (p >> 1) * V2_TILE_BYTES, *lt = l + (LO ? (p >> 1) * V2_TILE_BYTES : 0);
#pragma unroll
        for (int q = 0; q < 4; ++q) put8<RELU, LO>(ht, lt, row, (p & 1) * 4 + q, &v[8 * q]);
    }
}

// warp-level feed of hit-list / survivor-list indices: a chunk of 32 consecutive entries is claimed with ONE global atomic and
// handed out to the lanes that need a ray (uniform across the warp)
struct WarpFeed { int next, end; bool dry; };

}  // namespace

// the conditioning vectors of the clip as fp16 operand images: hi = rn16(x), lo = rn16(x - hi)  (64 values = 128 bytes per frame)
__global__ void k_cond_images(const float *__restrict__ cond /*[F,64]*/, int n, __half *__restrict__ hi, __half *__restrict__ lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = cond[i];
    const __half h = __float2half_rn(x);
    hi[i] = h;
    lo[i] = __float2half_rn(x - __half2float(h));
}

template <bool ROBUST, bool PROFILE>
__global__ void __launch_bounds__(Cfg<ROBUST>::NT, 1) k_head_v2(const __grid_constant__ HeadArgs a, const __grid_constant__ HeadV2Args t) {
    using C = Cfg<ROBUST>;
    constexpr int NSLOT = C::NSLOT, NSTAGE = C::NSTAGE, NTILE = C::NTILE, NWORK = C::NWORK;
    extern __shared__ __align__(1024) unsigned char smem_raw_[];
    unsigned char *sm = smem_raw_ + ((1024u - (smem_u32(smem_raw_) & 1023u)) & 1023u);
    Misc &ms = *reinterpret_cast<Misc *>(sm + C::OFF_MISC);
    unsigned char *wring = sm + C::OFF_W, *res = sm + C::OFF_RES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---------------- one-time setup ----------------
    MarchConst mc = a.mc;
    float occ_lo[3] = {0.f, 0.f, 0.f}, occ_hi[3] = {0.f, 0.f, 0.f};
    const bool have_box = setup_occupancy(a, mc, occ_lo, occ_hi);
    if (a.coarse_bits && a.coarse_words <= HEAD_COARSE_WORDS) {
        for (int i = tid; i < a.coarse_words; i += C::NT) ms.coarse[i] = a.coarse_bits[i];
        mc.coarse = ms.coarse;
    } else {
        mc.coarse = a.coarse_bits;
    }
    if (warp == NWORK) tmem_alloc(&ms.tmem_base, 512);
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(&ms.a_ready[i], 4); mbar_init(&ms.d_ready[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&ms.w_full[i], 1); mbar_init(&ms.w_free[i], 1); }
        mbar_init(&ms.res_full, 1);
        mbar_init(&ms.batch_bar, NWORK);
        mbar_fence_init();
        ms.stop = 0;
    }
    if (tid < 128) ms.cbias[tid] = a.narrow[7 * 128 + tid];
    if (tid < 32) { ms.flag[0][tid & 15] = 0; ms.flag[1][tid & 15] = 0; }
    stage_level_meta(a.pos_gm, ROBUST ? t.pos_step : nullptr, ms.lvl[0], tid);
    stage_level_meta(a.amb_gm, nullptr, ms.lvl[1], tid);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = ms.tmem_base;
    const int total = (a.pass == 1) ? *a.n_hits : *a.n_survivors;

    if (warp < NWORK) {
        // =====================================================================================================
        // ROW OWNER: thread `row` of slot `s`
        //
        // Software-pipelined over the sample sequence of its rays: `cu` is the sample being shaded in this batch, `sl` already
        // holds the NEXT one.  What happens to the ray after the current sample -- it dies because T < T_thresh (T uses the
        // weights BEFORE this sample, raymarching.cu:1000-1004), it reaches its sample cap, or marching runs off the volume --
        // does not depend on the sample's own sigma, so the march to the next sample, the adoption of a new ray and the gather
        // of the next sample's position features all run in the shadow of the current batch's MMAs.
        // =====================================================================================================
        const int s = warp >> 2, row = tid & (TM - 1);
        unsigned char *h = sm + s * C::SLOT_BYTES, *l = h + 2 * V2_TILE_BYTES, *sht = h + (ROBUST ? 4 : 2) * V2_TILE_BYTES;
        const uint32_t taddr = tmem + ((uint32_t)(warp & 3) * 32u << 16) + (uint32_t)s * D_COLS;
        const bool smooth_p = a.pos_gm.interp == 1, smooth_a = a.amb_gm.interp == 1;
        Slot sl;
        sl.active = false;
        sl.gid = 0; sl.frame = 0; sl.nsamp = 0; sl.cap = 0;
        sl.px = sl.py = sl.pz = 0.f; sl.dt = 0.f; sl.t = 0.f; sl.ws = 0.f; sl.depth = 0.f; sl.r = sl.gch = sl.b = 0.f;
        sl.near = sl.far = sl.far_m = 0.f;
        ray_geom_init(sl.g, 0.f, 0.f, 0.f, 1.f, 1.f, 1.f);
        WarpFeed wf{0, 0, false};
        uint32_t dpar = 0, bpar = 0;
        bool newray = false;      // sl holds a freshly adopted ray whose SH operand row is not written yet
        // the sample being shaded: accumulators BEFORE it, its dt / post-sample t, and what happens to its ray afterwards
        struct { bool valid; int fate, D, gid, frame, nsamp; float dt, t_post, ws, depth, r, g, b, near, far; } cu;
        cu.valid = false; cu.fate = 0; cu.D = 0; cu.gid = 0; cu.frame = 0; cu.nsamp = 0;
        cu.dt = cu.t_post = cu.ws = cu.depth = cu.r = cu.g = cu.b = cu.near = cu.far = 0.f;
        uint32_t feat[16];        // position features of the current sample (packed fp16 pairs), prefetched during the previous batch
        uint32_t featl[ROBUST ? 16 : 1];

#define ARRIVE_A()                                   \
    do {                                             \
        fence_async_smem();                          \
        fence_before_sync();                         \
        __syncwarp();                                \
        if (lane == 0) mbar_arrive(&ms.a_ready[s]);  \
    } while (0)
#define WAIT_D()                                     \
    do {                                             \
        mbar_wait(&ms.d_ready[s], dpar);             \
        dpar ^= 1u;                                  \
        fence_after_sync();                          \
    } while (0)
        // optional phase stamps (thread 0 = row 0 of slot 0 of every CTA): cycles between consecutive PH() points
        long long ph_last = clock64();
#define PH(i)                                                                              \
    if (PROFILE && a.phase_cycles && tid == 0) {                                                      \
        const long long now_ = clock64();                                                  \
        atomicAdd(a.phase_cycles + (i), (unsigned long long)(now_ - ph_last));             \
        ph_last = now_;                                                                    \
    }

        // refill dead lanes of `sl` from the warp's feed (all lanes of the warp, converged)
        auto refill = [&]() {
#pragma unroll 1
            for (int tries = 0; tries < 4; ++tries) {
                const unsigned need = __ballot_sync(0xffffffffu, !sl.active);
                if (!need) break;
                if (wf.next >= wf.end) {
                    if (wf.dry) break;
                    int base = 0;
                    if (lane == 0) base = atomicAdd(a.cursor, 32);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (base >= total) { wf.dry = true; break; }
                    wf.next = base;
                    wf.end = min(base + 32, total);
                }
                const int rank = __popc(need & ((1u << lane) - 1u)), avail = wf.end - wf.next;
                const bool take = !sl.active && rank < avail;
                const int w = wf.next + rank;
                wf.next += min(__popc(need), avail);
                if (take) {
                    bool live;
                    if (a.pass == 1) {
                        // k_ray_setup already set this ray up and marched it to its first sample: adopt the 64-byte record
                        HitRecord hr;
                        uint4 *q = reinterpret_cast<uint4 *>(&hr);
                        ldg256(a.hits + 4 * (size_t)w, q[0], q[1]);
                        ldg256(a.hits + 4 * (size_t)w + 2, q[2], q[3]);
                        sl.gid = hr.gid;
                        sl.frame = hr.gid / a.n_rays;
                        ray_geom_init(sl.g, hr.ox, hr.oy, hr.oz, hr.dx, hr.dy, hr.dz);
                        sl.near = hr.near; sl.far = hr.far; sl.far_m = hr.far_m;
                        sl.t = hr.t; sl.px = hr.px; sl.py = hr.py; sl.pz = hr.pz; sl.dt = hr.dt;
                        sl.ws = 0.f; sl.depth = 0.f; sl.r = sl.gch = sl.b = 0.f;
                        sl.nsamp = 0; sl.cap = a.max_steps;
                        live = true;
                    } else {
                        const int gid = a.survivors[w];
                        sl.gid = gid;
                        sl.frame = gid / a.n_rays;
                        load_ray(a, sl.frame, gid - sl.frame * a.n_rays, sl.g);
                        near_far(sl.g, a.aabb, a.min_near, sl.near, sl.far);
                        const size_t g = (size_t)gid;
                        sl.t = a.rays_t[g]; sl.ws = a.wsum[g]; sl.depth = a.depth[g];
                        sl.r = a.image[3 * g]; sl.gch = a.image[3 * g + 1]; sl.b = a.image[3 * g + 2];
                        sl.nsamp = a.max_steps; sl.cap = a.B_total[sl.frame];
                        (void)may_hit_occupied(have_box, occ_lo, occ_hi, sl.g, sl.near, sl.far, sl.far_m);
                        live = sl.nsamp < sl.cap && march_next(mc, sl.g, sl.far_m, sl.t, sl.px, sl.py, sl.pz, sl.dt);
                        if (!live) finalize_ray(a, sl, true);
                    }
                    sl.active = live;
                    newray = live;
                }
                __syncwarp();
            }
            // the records the warp will adopt next (64 B each, read once, from L2 or DRAM): pull them towards L1 a batch ahead
            if (a.pass == 1 && wf.next + lane < wf.end)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a.hits + 4 * (size_t)(wf.next + lane)));
        };
        // eight levels [8 * half, 8 * half + 8) of the position grid at the NEXT sample -> feat / featl
        auto prefetch_pos = [&](int half) {
            float f[16];
            if (sl.active) {
                const float inv2b = 2.0f * mc.bound;
                const float u = __fdiv_rn(__fadd_rn(sl.px, mc.bound), inv2b), vv = __fdiv_rn(__fadd_rn(sl.py, mc.bound), inv2b),
                            w = __fdiv_rn(__fadd_rn(sl.pz, mc.bound), inv2b);
                lookup8<ROBUST ? OCT_I16 : OCT_F16>(&ms.lvl[0][half * 16], a.pos_gm.align_off, smooth_p, a.pos_octs, u, vv, w, f);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t hi = pack2<false>(f[2 * i], f[2 * i + 1]);
                feat[half * 8 + i] = hi;
                if (ROBUST) {
                    const float2 b = unpack_h2(hi);
                    featl[half * 8 + i] = pack2<false>(f[2 * i] - b.x, f[2 * i + 1] - b.y);
                }
            }
        };
        // make `cu` the sample `sl` points at
        auto rotate = [&]() {
            cu.valid = sl.active;
            cu.dt = sl.dt; cu.t_post = sl.t;
            cu.ws = sl.ws; cu.depth = sl.depth; cu.r = sl.r; cu.g = sl.gch; cu.b = sl.b;
            cu.gid = sl.gid; cu.frame = sl.frame; cu.near = sl.near; cu.far = sl.far; cu.nsamp = sl.nsamp;
        };

        // ---------------- prologue: first rays, first samples, their position features ----------------
        refill();
        prefetch_pos(0);
        prefetch_pos(1);
        rotate();

        for (;;) {
            const bool valid = cu.valid;
            {   // batch accounting: this warp's "live rows" flag, then the split-phase batch barrier
                const unsigned any = __ballot_sync(0xffffffffu, valid);
                if (lane == 0) {
                    ms.flag[bpar][warp] = any ? 1 : 0;
                    mbar_arrive(&ms.batch_bar);
                }
            }
            // ---------------- ambient-net input: h0 k[0,32) <- position features, k[32,64) | h1 k[0,32) <- conditioning ----------------
            if (newray && valid) {   // SH(dir) is constant along the ray: its K16 operand row is written once per ray (shencoder.cu:43-68)
                float shv[16];
                sh4(sl.g.dx, sl.g.dy, sl.g.dz, shv);
                store_chunk<false, false>(sht, sht, k16_off(row, 0), &shv[0]);
                store_chunk<false, false>(sht, sht, k16_off(row, 1), &shv[8]);
                newray = false;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                *reinterpret_cast<uint4 *>(h + sw128_off(row, c)) = make_uint4(feat[4 * c], feat[4 * c + 1], feat[4 * c + 2], feat[4 * c + 3]);
                if (ROBUST) *reinterpret_cast<uint4 *>(l + sw128_off(row, c)) = make_uint4(featl[4 * c], featl[4 * c + 1], featl[4 * c + 2], featl[4 * c + 3]);
            }
            {   // the frame's conditioning vector as pre-rounded fp16 images (k_cond_images): 8 x 16 bytes straight into the operand tiles
                const uint4 *ch = reinterpret_cast<const uint4 *>(t.cond_hi) + (size_t)cu.frame * 8;
                const uint4 *cl = reinterpret_cast<const uint4 *>(t.cond_lo) + (size_t)cu.frame * 8;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t off = (c < 4) ? sw128_off(row, 4 + c) : (uint32_t)V2_TILE_BYTES + sw128_off(row, c - 4);
                    *reinterpret_cast<uint4 *>(h + off) = __ldg(ch + c);
                    if (ROBUST) *reinterpret_cast<uint4 *>(l + off) = __ldg(cl + c);
                }
            }
            PH(0)   // operand rows of ambient L0
            ARRIVE_A();   // -> ambient L0

            // ---------------- in the shadow of the ambient net: fate of the ray, march to the next sample, refill ----------------
            if (valid) {
                const int ns = cu.nsamp + 1;                 // samples composited once this one is
                const float T = 1.0f - cu.ws;
                if (T < a.T_thresh) { cu.fate = 1; cu.D = ns; }
                else if (ns >= sl.cap) cu.fate = 2;
                else if (!march_next(mc, sl.g, sl.far_m, sl.t, sl.px, sl.py, sl.pz, sl.dt)) { cu.fate = 1; cu.D = ns + 1; }
                else { cu.fate = 0; sl.nsamp = ns; }
                if (cu.fate) sl.active = false;
            }
            __syncwarp();
            refill();
            PH(1)   // fate + march + refill

            // ---------------- ambient net 96 -> 128 -> 128 -> 3 (radnerf.py:121) ----------------
            WAIT_D();
            PH(2)   // wait: ambient L0
            epilogue_row<true, 0, ROBUST>(taddr, h, l, row, nullptr);
            ARRIVE_A();   // -> ambient L1
            PH(3)   // epilogue: ambient L0
            WAIT_D();
            PH(4)   // wait: ambient L1
            epilogue_row<true, 0, ROBUST>(taddr, h, l, row, nullptr);
            ARRIVE_A();   // -> ambient out (N = 16)
            PH(5)   // epilogue: ambient L1
            WAIT_D();
            PH(6)   // wait: ambient out
            float amb0, amb1, amb2;
            {
                float o[16];
                tmem_ld16(taddr + 128u, o);
                wait_ld();
                amb0 = tanhf(o[0]); amb1 = tanhf(o[1]); amb2 = tanhf(o[2]);
            }
            // ---------------- sigma-net input: h0 k[0,32) <- position features (registers), k[32,64) <- ambient grid ----------------
            {
                float u = 0.f, vv = 0.f, w = 0.f;
                if (valid) {   // GridEncoder.forward with bound = 1: (x + 1) / 2
                    u = __fdiv_rn(__fadd_rn(amb0, 1.0f), 2.0f);
                    vv = __fdiv_rn(__fadd_rn(amb1, 1.0f), 2.0f);
                    w = __fdiv_rn(__fadd_rn(amb2, 1.0f), 2.0f);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4 *>(h + sw128_off(row, c)) = make_uint4(feat[4 * c], feat[4 * c + 1], feat[4 * c + 2], feat[4 * c + 3]);
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    float f[16];
                    if (valid) {
                        lookup8<OCT_F16>(&ms.lvl[1][half * 16], a.amb_gm.align_off, smooth_a, a.amb_octs, u, vv, w, f);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) f[i] = 0.f;
                    }
                    put8<false, false>(h, h, row, 4 + 2 * half, &f[0]);
                    put8<false, false>(h, h, row, 4 + 2 * half + 1, &f[8]);
                }
            }
            PH(7)   // tanh + ambient gather
            ARRIVE_A();   // -> sigma L0

            // ---------------- sigma net 64 -> 128 -> 128 -> (128 geo + sigma) (radnerf.py:127-130); in its shadow: the NEXT sample's
            //                  position features (the current ones are in the operand tile now) ----------------
            prefetch_pos(0);
            PH(8)   // prefetch: position levels 0-7 of the next sample
            WAIT_D();
            PH(9)   // wait: sigma L0
            epilogue_row<true, 0, false>(taddr, h, h, row, nullptr);
            ARRIVE_A();   // -> sigma L1
            PH(10)  // epilogue: sigma L0
            prefetch_pos(1);
            PH(11)  // prefetch: position levels 8-15 of the next sample
            WAIT_D();
            PH(12)  // wait: sigma L1
            epilogue_row<true, 0, false>(taddr, h, h, row, nullptr);
            ARRIVE_A();   // -> sigma L2 (geo rows + the sigma row as an N = 16 group)
            PH(13)  // epilogue: sigma L1
            WAIT_D();
            PH(14)  // wait: sigma L2
            float sigma;
            {
                float o[16];
                tmem_ld16(taddr + 128u, o);
                wait_ld();
                sigma = a.density_scale * expf(o[0]);   // trunc_exp forward (utils.py:36-49) * density_scale (renderer.py:365)
            }
            epilogue_row<false, 0, false>(taddr, h, h, row, nullptr);   // geo features (no activation) -> color-net input
            ARRIVE_A();   // -> color L0 (geo K = 128 + SH K = 16)
            PH(15)  // epilogue: sigma L2 (sigma + geo)

            // ---------------- color net (16 SH + 128 geo [+ folded individual code]) -> 128 -> 3 (radnerf.py:137-141) ----------------
            WAIT_D();
            PH(16)  // wait: color L0
            epilogue_row<true, 2, false>(taddr, h, h, row, ms.cbias);
            ARRIVE_A();   // -> color out (N = 16)
            PH(17)  // epilogue: color L0
            WAIT_D();
            PH(18)  // wait: color out
            float cr, cg, cb;
            {
                float o[16];
                tmem_ld16(taddr + 128u, o);
                wait_ld();
                cr = 1.0f / (1.0f + expf(-o[0]));
                cg = 1.0f / (1.0f + expf(-o[1]));
                cb = 1.0f / (1.0f + expf(-o[2]));
            }
            fence_before_sync();   // the next batch's MMAs overwrite these TMEM columns: order the loads above before the next arrive

            // ---------------- composite the sample (raymarching.cu:978-1006) and retire / suspend / continue its ray ----------------
            if (valid) {
                const float alpha = 1.0f - expf(-sigma * cu.dt);
                const float T = 1.0f - cu.ws;
                const float wgt = alpha * T;
                cu.ws += wgt;
                cu.depth += wgt * cu.t_post;   // deltas[1]: t after the sample
                cu.r += wgt * cr;
                cu.g += wgt * cg;
                cu.b += wgt * cb;
                if (a.valid_samples) warp_agg_add(a.valid_samples, cu.frame, 1);
                if (cu.fate == 0) {            // same ray goes on: `sl` (already at its next sample) takes the accumulators
                    sl.ws = cu.ws; sl.depth = cu.depth; sl.r = cu.r; sl.gch = cu.g; sl.b = cu.b;
                } else {
                    const size_t g = (size_t)cu.gid;
                    const bool normalise = cu.fate == 1 || a.pass != 1;   // a suspended pass-1 ray keeps its raw depth: pass 2 goes on accumulating
                    a.image[3 * g] = cu.r; a.image[3 * g + 1] = cu.g; a.image[3 * g + 2] = cu.b;
                    a.wsum[g] = cu.ws;
                    // renderer.py:394: depth = clamp(depth - nears, min=0) / (fars - nears)
                    a.depth[g] = normalise ? __fdiv_rn(fmaxf(__fsub_rn(cu.depth, cu.near), 0.f), __fsub_rn(cu.far, cu.near)) : cu.depth;
                    if (cu.fate == 1) {
                        if (a.pass == 1) warp_agg_add(a.hist, cu.frame * (a.max_steps + 2) + cu.D, 1);
                    } else if (a.pass == 1) {
                        a.rays_t[g] = cu.t_post;
                        cg::coalesced_group grp = cg::coalesced_threads();
                        int base = 0;
                        if (grp.thread_rank() == 0) base = atomicAdd(a.n_survivors, (int)grp.size());
                        base = grp.shfl(base, 0);
                        a.survivors[base + grp.thread_rank()] = cu.gid;
                    }
                }
            }
            __syncwarp();
            rotate();

            // ---------------- end of batch: was anything alive in it? ----------------
            mbar_wait(&ms.batch_bar, bpar);
            int any = 0;
#pragma unroll
            for (int i = 0; i < NWORK; ++i) any |= ms.flag[bpar][i];
            bpar ^= 1u;
            __syncwarp();
            PH(19)  // sigmoid + composite + end-of-batch check
            if (PROFILE && a.phase_cycles && tid == 0) atomicAdd(a.phase_cycles + 31, 1ull);   // batches
            if (!any) break;
        }
#undef PH
#undef ARRIVE_A
#undef WAIT_D
    } else if (warp == NWORK) {
        // =====================================================================================================
        // ISSUE WARP: layer-major, slot-minor; all lanes converged, one elected lane issues
        // =====================================================================================================
        const uint32_t idesc128 = make_idesc(0, 128), idesc16 = make_idesc(0, 16);
        const uint32_t res_u = smem_u32(res), ring_u = smem_u32(wring), slot_u = smem_u32(sm);
        uint32_t apar = 0;          // all slots arrive the same number of times: one parity bit serves them all
        uint32_t wc = 0;            // streamed tiles consumed so far
        uint32_t bpar = 0;
        mbar_wait(&ms.res_full, 0);
        fence_after_sync();

        auto h_of = [&](int s, int i) -> uint32_t { return slot_u + (uint32_t)(s * C::SLOT_BYTES + i * V2_TILE_BYTES); };
        auto l_of = [&](int s, int i) -> uint32_t { return slot_u + (uint32_t)(s * C::SLOT_BYTES + (2 + i) * V2_TILE_BYTES); };
        auto sh_of = [&](int s) -> uint32_t { return slot_u + (uint32_t)(s * C::SLOT_BYTES + (ROBUST ? 4 : 2) * V2_TILE_BYTES); };
        auto d_of = [&](int s) -> uint32_t { return tmem + (uint32_t)s * D_COLS; };
        // wait for streamed tile wc + j, return its shared-memory address
        auto tile = [&](uint32_t j) -> uint32_t {
            const uint32_t i = wc + j, stage = i % NSTAGE;
            mbar_wait(&ms.w_full[stage], (i / NSTAGE) & 1u);
            return ring_u + stage * V2_TILE_BYTES;
        };
        auto release = [&](uint32_t n) {   // the n tiles just used are free once every MMA issued so far has completed
            if (elect_one()) {
                for (uint32_t j = 0; j < n; ++j) mma_commit(&ms.w_free[(wc + j) % NSTAGE]);
            }
            __syncwarp();
            wc += n;
        };
        // one 128-wide ambient layer in the split arithmetic: two K tiles, each as (hi, lo) weight images shared by all slots;
        // D = Al*Wh + Ah*Wl + Ah*Wh, small terms first
        auto robust_layer = [&](int ks0, int ks1) {
#pragma unroll
            for (int kt = 0; kt < 2; ++kt) {
                const uint32_t wh = tile(0), wl = tile(1);
                const int ks = kt ? ks1 : ks0;
                fence_after_sync();
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    if (kt == 0) { mbar_wait(&ms.a_ready[s], apar); fence_after_sync(); }
                    if (elect_one()) {
                        mma_sw(d_of(s), l_of(s, kt), wh, ks, idesc128, kt ? 1u : 0u);
                        mma_sw(d_of(s), h_of(s, kt), wl, ks, idesc128, 1u);
                        mma_sw(d_of(s), h_of(s, kt), wh, ks, idesc128, 1u);
                        if (kt == 1) mma_commit(&ms.d_ready[s]);
                    }
                    __syncwarp();
                }
                release(2);
            }
        };
        // Slots of a layer are served in index order (blocking waits) or -- `relaxed` -- in the order their owners arrive: a sweep of
        // non-blocking tests serves whoever is ready; only when nobody is does the warp block on the first slot still pending.
        const bool relaxed = (t.flags & 1u) != 0u;
#define FOR_SLOTS                                                                                                     \
    for (uint32_t pend = (1u << NSLOT) - 1u, force = relaxed ? 0u : 1u, got = 0u; pend; force = (relaxed && got) ? 0u : 1u, got = 0u) \
        _Pragma("unroll") for (int s = 0; s < NSLOT; ++s)
#define SLOT_BEGIN(sv)                                                         \
    if (!((pend >> (sv)) & 1u)) continue;                                      \
    if (force) { mbar_wait(&ms.a_ready[sv], apar); if (relaxed) force = 0u; }  \
    else if (!mbar_test(&ms.a_ready[sv], apar)) continue;                      \
    pend &= ~(1u << (sv)); got = 1u;                                           \
    fence_after_sync();                                                        \
    if (elect_one()) {
#define SLOT_END(sv)                              \
        mma_commit(&ms.d_ready[sv]);              \
    }                                             \
    __syncwarp();

        for (;;) {
            // ---- ambient L0: K = 96 = 32 position features + 64 conditioning values (h0 k[0,64), h1 k[0,32)) ----
            if (ROBUST) {
                robust_layer(4, 2);
            } else {
                const uint32_t w0 = tile(0), w1 = tile(1);
                fence_after_sync();
                FOR_SLOTS {
                    SLOT_BEGIN(s)
                    mma_sw(d_of(s), h_of(s, 0), w0, 4, idesc128, 0u);
                    mma_sw(d_of(s), h_of(s, 1), w1, 2, idesc128, 1u);
                    SLOT_END(s)
                }
                release(2);
            }
            apar ^= 1u;
            // everyone has arrived for this batch: is anything alive in it?  (read now, acted upon at the end of the batch)
            mbar_wait(&ms.batch_bar, bpar);
            int any = 0;
#pragma unroll
            for (int i = 0; i < NWORK; ++i) any |= ms.flag[bpar][i];
            bpar ^= 1u;
            // ---- ambient L1: K = 128 ----
            if (ROBUST) {
                robust_layer(4, 4);
            } else {
                const uint32_t w0 = tile(0), w1 = tile(1);
                fence_after_sync();
                FOR_SLOTS {
                    SLOT_BEGIN(s)
                    mma_sw(d_of(s), h_of(s, 0), w0, 4, idesc128, 0u);
                    mma_sw(d_of(s), h_of(s, 1), w1, 4, idesc128, 1u);
                    SLOT_END(s)
                }
                release(2);
            }
            apar ^= 1u;
            // ---- ambient out: N = 16 group from the resident tiles, into columns 128..143 ----
            FOR_SLOTS {
                SLOT_BEGIN(s)
                if (ROBUST) {
                    mma_sw(d_of(s) + 128u, l_of(s, 0), res_u + V2_RES_AMBN_HI, 4, idesc16, 0u);
                    mma_sw(d_of(s) + 128u, l_of(s, 1), res_u + V2_RES_AMBN_HI + 2048, 4, idesc16, 1u);
                    mma_sw(d_of(s) + 128u, h_of(s, 0), res_u + V2_RES_AMBN_LO, 4, idesc16, 1u);
                    mma_sw(d_of(s) + 128u, h_of(s, 1), res_u + V2_RES_AMBN_LO + 2048, 4, idesc16, 1u);
                    mma_sw(d_of(s) + 128u, h_of(s, 0), res_u + V2_RES_AMBN_HI, 4, idesc16, 1u);
                    mma_sw(d_of(s) + 128u, h_of(s, 1), res_u + V2_RES_AMBN_HI + 2048, 4, idesc16, 1u);
                } else {
                    mma_sw(d_of(s) + 128u, h_of(s, 0), res_u + V2_RES_AMBN_HI, 4, idesc16, 0u);
                    mma_sw(d_of(s) + 128u, h_of(s, 1), res_u + V2_RES_AMBN_HI + 2048, 4, idesc16, 1u);
                }
                SLOT_END(s)
            }
            apar ^= 1u;
            // ---- sigma L0: K = 64 ----
            {
                const uint32_t w0 = tile(0);
                fence_after_sync();
                FOR_SLOTS {
                    SLOT_BEGIN(s)
                    mma_sw(d_of(s), h_of(s, 0), w0, 4, idesc128, 0u);
                    SLOT_END(s)
                }
                release(1);
                apar ^= 1u;
            }
            // ---- sigma L1: K = 128 ----
            {
                const uint32_t w0 = tile(0), w1 = tile(1);
                fence_after_sync();
                FOR_SLOTS {
                    SLOT_BEGIN(s)
                    mma_sw(d_of(s), h_of(s, 0), w0, 4, idesc128, 0u);
                    mma_sw(d_of(s), h_of(s, 1), w1, 4, idesc128, 1u);
                    SLOT_END(s)
                }
                release(2);
                apar ^= 1u;
            }
            // ---- sigma L2: 128 geo rows + the sigma row (resident N = 16 tile -> column 128) ----
            {
                const uint32_t w0 = tile(0), w1 = tile(1);
                fence_after_sync();
                FOR_SLOTS {
                    SLOT_BEGIN(s)
                    mma_sw(d_of(s), h_of(s, 0), w0, 4, idesc128, 0u);
                    mma_sw(d_of(s), h_of(s, 1), w1, 4, idesc128, 1u);
                    mma_sw(d_of(s) + 128u, h_of(s, 0), res_u + V2_RES_SIGROW, 4, idesc16, 0u);
                    mma_sw(d_of(s) + 128u, h_of(s, 1), res_u + V2_RES_SIGROW + 2048, 4, idesc16, 1u);
                    SLOT_END(s)
                }
                release(2);
                apar ^= 1u;
            }
            // ---- color L0: geo K = 128 (streamed) + SH K = 16 (resident no-swizzle tile against the slot's SH operand tile) ----
            {
                const uint32_t w0 = tile(0), w1 = tile(1);
                fence_after_sync();
                FOR_SLOTS {
                    SLOT_BEGIN(s)
                    mma_sw(d_of(s), h_of(s, 0), w0, 4, idesc128, 0u);
                    mma_sw(d_of(s), h_of(s, 1), w1, 4, idesc128, 1u);
                    mma_f16(d_of(s), desc_k16(sh_of(s)), desc_k16(res_u + V2_RES_COLSH), idesc128, 1u);
                    SLOT_END(s)
                }
                release(2);
                apar ^= 1u;
            }
            // ---- color out: N = 16 group ----
            FOR_SLOTS {
                SLOT_BEGIN(s)
                mma_sw(d_of(s) + 128u, h_of(s, 0), res_u + V2_RES_COLN, 4, idesc16, 0u);
                mma_sw(d_of(s) + 128u, h_of(s, 1), res_u + V2_RES_COLN + 2048, 4, idesc16, 1u);
                SLOT_END(s)
            }
            apar ^= 1u;
            if (!any) break;
        }
#undef FOR_SLOTS
#undef SLOT_BEGIN
#undef SLOT_END
        // stop the loader: the NSTAGE tiles it has prefetched for a batch that never comes must land before the CTA may exit
        ms.stop = 1;
        __threadfence_block();
        for (uint32_t j = 0; j < (uint32_t)NSTAGE; ++j) {
            const uint32_t i = wc + j, stage = i % NSTAGE;
            mbar_wait(&ms.w_full[stage], (i / NSTAGE) & 1u);
            if (lane == 0) mbar_arrive(&ms.w_free[stage]);
            __syncwarp();
        }
    } else {
        // =====================================================================================================
        // LOADER WARP: resident tiles once, then the weight stream, NSTAGE tiles ahead of the issue warp
        // =====================================================================================================
        if (elect_one()) {
            mbar_expect_tx(&ms.res_full, V2_RES_BYTES);
            bulk_g2s(res, t.w_res, V2_RES_BYTES, &ms.res_full);
        }
        __syncwarp();
        for (uint32_t i = 0;; ++i) {
            const uint32_t stage = i % NSTAGE;
            if (i >= (uint32_t)NSTAGE) mbar_wait(&ms.w_free[stage], ((i / NSTAGE) - 1u) & 1u);
            if (ms.stop) break;
            if (elect_one()) {
                mbar_expect_tx(&ms.w_full[stage], V2_TILE_BYTES);
                bulk_g2s(wring + stage * V2_TILE_BYTES, t.w_stream + (size_t)(i % NTILE) * V2_TILE_BYTES, V2_TILE_BYTES, &ms.w_full[stage]);
            }
            __syncwarp();
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == NWORK) tmem_dealloc(tmem, 512);
}

size_t head_v2_smem_bytes(bool robust) { return robust ? smem_bytes<true>() : smem_bytes<false>(); }

cudaError_t launch_cond_images(const float *cond_feat, int n_frames, void *cond_hi, void *cond_lo, cudaStream_t st) {
    const int n = n_frames * 64;
    k_cond_images<<<(n + 255) / 256, 256, 0, st>>>(cond_feat, n, (__half *)cond_hi, (__half *)cond_lo);
    return cudaGetLastError();
}

cudaError_t launch_head_v2(const HeadArgs &a, const HeadV2Args &t_in, int precision, cudaStream_t st) {
    HeadV2Args t = t_in;
    // scheduling knob of the issue warp (bit 0: serve the slots of a layer in arrival order); GFPP_V2_RELAXED overrides the default for A/B runs
    t.flags = getenv("GFPP_V2_RELAXED") ? (atoi(getenv("GFPP_V2_RELAXED")) ? 1u : 0u) : V2_DEFAULT_FLAGS;
    const bool robust = precision == FP16_ROBUST;
    if (!robust && precision != FP16_X1) return cudaErrorInvalidValue;
    const int blocks = sm_count();   // persistent: one CTA per SM (512 TMEM columns and ~200 KB of shared memory each)
    const size_t smem = head_v2_smem_bytes(robust);
    cudaError_t e;
    // function attributes are per device: set on every launch.  The phase-stamped instantiation only runs under tools/phase_breakdown.py.
#define GO(RB, PF)                                                                                                   \
    e = cudaFuncSetAttribute(k_head_v2<RB, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
    if (e != cudaSuccess) return e;                                                                                  \
    k_head_v2<RB, PF><<<blocks, Cfg<RB>::NT, smem, st>>>(a, t);
    if (robust) { if (a.phase_cycles) { GO(true, true) } else { GO(true, false) } }
    else { if (a.phase_cycles) { GO(false, true) } else { GO(false, false) } }
#undef GO
    return cudaGetLastError();
}

}  // namespace gfpp
