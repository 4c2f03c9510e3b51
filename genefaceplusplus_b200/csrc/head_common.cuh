// head_common.cuh -- ray-slot state and helpers shared by the two fused head kernels
// (head_kernel.cu: fp32 FFMA MLP; head_tc_kernel.cu: tcgen05 tensor-core MLP).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "head_kernel.cuh"

namespace gfpp {
namespace headc {

namespace cg = cooperative_groups;

// add `val` to base[key] once per distinct key in the warp's active lanes
__device__ __forceinline__ void warp_agg_add(int *base, int key, int val) {
    const unsigned act = __activemask();
    const unsigned grp = __match_any_sync(act, key);
    const int sum = __reduce_add_sync(grp, val);
    if ((int)(__ffs(grp) - 1) == (int)(threadIdx.x & 31)) atomicAdd(base + key, sum);
}

struct Slot {
    RayGeom g;
    float t, near, far, ws, depth, r, gch, b;
    float far_m;   // marching bound: min(far, exit of the padded occupied box) -- nothing is occupied beyond it
    float px, py, pz, dt;  // pending sample
    int gid, frame, nsamp, cap;
    bool active;
};

__device__ __forceinline__ void load_ray(const HeadArgs &a, int frame, int ray, RayGeom &g) {
    if (a.rays_o) {
        const size_t o = ((size_t)frame * a.n_rays + ray) * 3;
        ray_geom_init(g, a.rays_o[o], a.rays_o[o + 1], a.rays_o[o + 2], a.rays_d[o], a.rays_d[o + 1], a.rays_d[o + 2]);
    } else {
        // get_rays (modules/radnerfs/utils.py:302-360): pixel centre, normalise, rotate by c2w[:3,:3]
        const float *P = a.poses + (size_t)frame * 16;
        const int row = ray / a.img_w, col = ray - row * a.img_w;
        const float xs = __fdiv_rn(__fsub_rn((float)col + 0.5f, a.cx), a.fx);
        const float ys = __fdiv_rn(__fsub_rn((float)row + 0.5f, a.cy), a.fy);
        const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(xs, xs), __fmul_rn(ys, ys)), 1.0f));
        const float dxc = __fdiv_rn(xs, nrm), dyc = __fdiv_rn(ys, nrm), dzc = __fdiv_rn(1.0f, nrm);
        float d[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            d[k] = __fadd_rn(__fadd_rn(__fmul_rn(dxc, P[4 * k]), __fmul_rn(dyc, P[4 * k + 1])), __fmul_rn(dzc, P[4 * k + 2]));
        ray_geom_init(g, P[3], P[7], P[11], d[0], d[1], d[2]);
    }
}

// conservative: can the segment [near, far] of the ray touch the (one-cell padded) box of occupied voxels?
// On return `far_m` is the parameter beyond which the ray cannot meet an occupied voxel (<= far): the marcher may stop
// there, which removes the long ALU-only walks of rays that have left the object.
__device__ __forceinline__ bool may_hit_occupied(bool have_box, const float (&occ_lo)[3], const float (&occ_hi)[3],
                                                 const RayGeom &g, float near, float far, float &far_m) {
    far_m = far;
    if (!have_box) return true;
    float t0 = near, t1 = far;
    const float o[3] = {g.ox, g.oy, g.oz}, rd[3] = {g.rdx, g.rdy, g.rdz}, d[3] = {g.dx, g.dy, g.dz};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (d[k] == 0.f) {
            if (o[k] < occ_lo[k] || o[k] > occ_hi[k]) return false;
            continue;
        }
        float ta = (occ_lo[k] - o[k]) * rd[k], tb = (occ_hi[k] - o[k]) * rd[k];
        if (ta > tb) { const float s = ta; ta = tb; tb = s; }
        t0 = fmaxf(t0, ta);
        t1 = fminf(t1, tb);
    }
    if (t0 <= t1) far_m = fminf(far, t1);
    return t0 <= t1;
}

__device__ __forceinline__ void finalize_ray(const HeadArgs &a, const Slot &s, bool normalise_depth) {
    const size_t g = (size_t)s.gid;
    a.image[3 * g] = s.r;
    a.image[3 * g + 1] = s.gch;
    a.image[3 * g + 2] = s.b;
    a.wsum[g] = s.ws;
    // renderer.py:394: depth = clamp(depth - nears, min=0) / (fars - nears)
    a.depth[g] = normalise_depth ? __fdiv_rn(fmaxf(__fsub_rn(s.depth, s.near), 0.f), __fsub_rn(s.far, s.near)) : s.depth;
}



// padded world-space box of the occupied voxels (cascade 0) from the device-side cell bounds; also installs the cell
// bounds into `mc`.  Rays whose [near, far] segment misses the box have no sample.
__device__ __forceinline__ bool setup_occupancy(const HeadArgs &a, MarchConst &mc, float (&occ_lo)[3], float (&occ_hi)[3]) {
    bool have_box = false;
    if (a.occ_bounds) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { mc.bb_lo[k] = a.occ_bounds[k]; mc.bb_hi[k] = a.occ_bounds[3 + k]; }
        if (a.use_occ_box) {
            have_box = true;
            const float mb = fminf(1.0f, mc.bound);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (mc.bb_hi[k] < mc.bb_lo[k]) { occ_lo[k] = 1e30f; occ_hi[k] = -1e30f; }  // nothing occupied
                else {
                    occ_lo[k] = ((float)(mc.bb_lo[k] - 1) * mc.rH * 2.0f - 1.0f) * mb;
                    occ_hi[k] = ((float)(mc.bb_hi[k] + 2) * mc.rH * 2.0f - 1.0f) * mb;
                }
            }
        }
    }
    return have_box;
}

// one out-of-line copy of the marcher per kernel (three call sites; the kernels must stay I-cache friendly)
// (results by value: reference parameters would force the marcher's loop state through local memory)
struct MarchOut {
    float t, x, y, z, dt;
    int ok;
};
static __device__ __noinline__ MarchOut march_next_out(const MarchConst &mc, const RayGeom &g, float far, float t0) {
    float t = t0, x = 0.f, y = 0.f, z = 0.f, dt = 0.f;
    const bool ok = march_next(mc, g, far, t, x, y, z, dt);
    MarchOut o;
    o.t = t; o.x = x; o.y = y; o.z = z; o.dt = dt; o.ok = ok ? 1 : 0;
    return o;
}
__device__ __forceinline__ bool march_next_nl(const MarchConst &mc, const RayGeom &g, float far, float &t, float &x, float &y, float &z,
                                              float &dt) {
#ifdef GFPP_OUTLINE_MARCH
    const MarchOut o = march_next_out(mc, g, far, t);
    t = o.t; x = o.x; y = o.y; z = o.z; dt = o.dt;
    return o.ok != 0;
#else
    return march_next(mc, g, far, t, x, y, z, dt);
#endif
}

// ---- partner prefetch --------------------------------------------------------------------------------------------
// Threads TM..2*TM-1 own no ray slot and used to idle while the owners composite.  Thread TM+i is the PARTNER of slot i:
// during the composite phase it fetches candidate rays from the CTA's work chunk and marches them (a bounded number of
// cell steps per round, resumable) until one reaches its first occupied sample, then parks {ray id, t_pre} in
// s.spare_*[i].  When slot i dies, its owner adopts the spare in O(1) instead of marching a new ray while the whole CTA
// waits at a barrier.  s.spare_gid codes: >= 0 ready, -1 empty (partner idle), -2 partner busy (will deliver).
struct Partner {
    RayGeom g;
    float far, far_m, t;
    int gid;
    int state;   // 0 idle, 1 marching
};
constexpr int kPartnerBudget = 24;   // cost units per round: 4 per fine-bitfield read, 1 per ALU-only cell step

template <class SmemT>
__device__ __forceinline__ void partner_step(const HeadArgs &a, SmemT &s, Partner &p, const MarchConst &mc, bool have_box,
                                             const float (&occ_lo)[3], const float (&occ_hi)[3], int tid) {
    constexpr int TM = HEAD_TM;
    if (tid < TM || a.pass != 1 || a.partner_budget <= 0) return;
    const int i = tid - TM;
    int budget = a.partner_budget;
    while (budget > 0) {
        if (p.state == 0) {
            if (s.spare_gid[i] != -1 || s.next >= s.end) break;      // spare still parked, or no work in the local chunk
            const int w = atomicAdd(&s.next, 1);
            if (w >= s.end) break;
            budget -= 2;
            p.gid = w;
            const int frame = w / a.n_rays, ray = w - frame * a.n_rays;
            load_ray(a, frame, ray, p.g);
            float near;
            near_far(p.g, a.aabb, a.min_near, near, p.far);
            p.t = near;
            if (!may_hit_occupied(have_box, occ_lo, occ_hi, p.g, near, p.far, p.far_m)) {
                // no sample at all: the ray dies at position 1 (delta == 0): zeros out, depth normalised like the reference
                Slot z;
                z.g = p.g; z.near = near; z.far = p.far; z.gid = w; z.frame = frame;
                z.ws = 0.f; z.depth = 0.f; z.r = z.gch = z.b = 0.f;
                finalize_ray(a, z, true);
                warp_agg_add(a.hist, frame * (a.max_steps + 2) + 1, 1);
                continue;
            }
            p.state = 1;
            s.spare_gid[i] = -2;
        }
        const int r = march_budget(mc, p.g, p.far_m, p.t, budget);
        if (r == 2) break;                                            // out of budget: resume next round
        if (r == 1) {                                                 // found: park it
            s.spare_t[i] = p.t;
            s.spare_gid[i] = p.gid;
            p.state = 0;
            break;
        }
        // exhausted without a sample
        {
            const int frame = p.gid / a.n_rays;
            Slot z;
            z.g = p.g; z.far = p.far; z.gid = p.gid; z.frame = frame;
            float nr, fr;
            near_far(p.g, a.aabb, a.min_near, nr, fr);
            z.near = nr;
            z.ws = 0.f; z.depth = 0.f; z.r = z.gch = z.b = 0.f;
            finalize_ray(a, z, true);
            warp_agg_add(a.hist, frame * (a.max_steps + 2) + 1, 1);
        }
        p.state = 0;
        s.spare_gid[i] = -1;
    }
}

constexpr int kFetchTries = 4;   // rays a thread may try per refill iteration (most candidates are cheap misses)
constexpr int kRefillIters = 2;

// Refill dead slots from the global work cursor, then publish the batch (valid flags, frame ids, sample positions).
// Returns the number of valid rows, or -1 when the CTA is out of work.  All threads of the CTA must call it.
template <class SmemT>
__device__ __forceinline__ int refill_and_publish(const HeadArgs &a, SmemT &s, Slot &sl, Partner &pt, const MarchConst &mc, bool have_box,
                                                  const float (&occ_lo)[3], const float (&occ_hi)[3], int total, int tid) {
    constexpr int TM = HEAD_TM;
    for (int it = 0; it < kRefillIters; ++it) {
        if (tid == 0 && s.next >= s.end && !s.done) {
            const int base = atomicAdd(a.cursor, TM);
            if (base >= total) { s.done = 1; }
            else { s.next = base; s.end = min(base + TM, total); }
        }
        __syncthreads();
        if (tid < TM && !sl.active && s.spare_gid[tid] >= 0) {
            // adopt the ray the partner thread pre-marched: O(1), no marching on the critical path
            const int gid = s.spare_gid[tid];
            sl.gid = gid;
            sl.frame = gid / a.n_rays;
            load_ray(a, sl.frame, gid - sl.frame * a.n_rays, sl.g);
            near_far(sl.g, a.aabb, a.min_near, sl.near, sl.far);
            (void)may_hit_occupied(have_box, occ_lo, occ_hi, sl.g, sl.near, sl.far, sl.far_m);
            sl.ws = 0.f; sl.depth = 0.f; sl.r = sl.gch = sl.b = 0.f;
            sl.nsamp = 0; sl.cap = a.max_steps;
            sample_at(mc, sl.g, s.spare_t[tid], sl.t, sl.px, sl.py, sl.pz, sl.dt);
            sl.active = true;
            s.spare_gid[tid] = -1;
        }
        if (tid < TM && !sl.active && s.spare_gid[tid] == -1) {       // -2: the partner is about to deliver, just wait
            for (int attempt = 0; attempt < kFetchTries && !sl.active && s.next < s.end; ++attempt) {
                const int w = atomicAdd(&s.next, 1);
                if (w >= s.end) break;
                int gid;
                float t_pre = 0.f;
                if (a.pass == 2) gid = a.survivors[w];
                else { const int2 hv = a.hits[w]; gid = hv.x; t_pre = __int_as_float(hv.y); }
                sl.gid = gid;
                sl.frame = gid / a.n_rays;
                const int ray = gid - sl.frame * a.n_rays;
                load_ray(a, sl.frame, ray, sl.g);
                near_far(sl.g, a.aabb, a.min_near, sl.near, sl.far);
                bool live;
                if (a.pass == 1) {
                    // k_ray_setup already marched this ray to its first sample: adopt it in O(1)
                    sl.ws = 0.f; sl.depth = 0.f; sl.r = sl.gch = sl.b = 0.f;
                    sl.nsamp = 0; sl.cap = a.max_steps;
                    (void)may_hit_occupied(have_box, occ_lo, occ_hi, sl.g, sl.near, sl.far, sl.far_m);
                    sample_at(mc, sl.g, t_pre, sl.t, sl.px, sl.py, sl.pz, sl.dt);
                    live = true;
                } else {
                    const size_t g = (size_t)gid;
                    sl.t = a.rays_t[g]; sl.ws = a.wsum[g]; sl.depth = a.depth[g];
                    sl.r = a.image[3 * g]; sl.gch = a.image[3 * g + 1]; sl.b = a.image[3 * g + 2];
                    sl.nsamp = a.max_steps; sl.cap = a.B_total[sl.frame];
                    (void)may_hit_occupied(have_box, occ_lo, occ_hi, sl.g, sl.near, sl.far, sl.far_m);
                    live = sl.nsamp < sl.cap && march_next_nl(mc, sl.g, sl.far_m, sl.t, sl.px, sl.py, sl.pz, sl.dt);
                    if (!live) finalize_ray(a, sl, true);
                }
                sl.active = live;
            }
        }
        // another iteration only pays off when slots are still empty AND the local chunk ran dry while work remains
        const int want_more = __syncthreads_or(tid < TM && !sl.active && s.spare_gid[tid] == -1 && s.next >= s.end && !s.done);
        if (!want_more) break;
    }
    if (tid < TM) {
        s.valid[tid] = sl.active ? 1 : 0;
        s.frame[tid] = sl.frame;
        s.sx[tid] = sl.px; s.sy[tid] = sl.py; s.sz[tid] = sl.pz;
    }
    const int n_valid = __syncthreads_count(tid < TM && sl.active);
    if (n_valid == 0) {
        // nothing to evaluate this round: finished only when the work is gone AND no partner still holds / marches a ray
        const int pending = __syncthreads_count((tid < TM && s.spare_gid[tid] != -1) || (tid >= TM && pt.state != 0));
        const bool out_of_work = s.done && s.next >= s.end && pending == 0;
        __syncthreads();  // thread 0 must not start the next refill (which rewrites next/end/done) before everyone has read them
        if (out_of_work) return -1;
        partner_step(a, s, pt, mc, have_box, occ_lo, occ_hi, tid);   // keep the prefetchers moving
        __syncthreads();  // ... and finished before thread 0 may hand out the next chunk (they read next/end)
        return 0;
    }
    return n_valid;
}

// Front-to-back compositing of the batch's sample in the owner thread (raymarching.cu:978-1006), termination test,
// then march to the ray's next sample (or retire the ray).
template <class SmemT>
__device__ __forceinline__ void composite_and_advance(const HeadArgs &a, SmemT &s, Slot &sl, Partner &pt, const MarchConst &mc,
                                                      bool have_box, const float (&occ_lo)[3], const float (&occ_hi)[3], int tid) {
    constexpr int TM = HEAD_TM;
    partner_step(a, s, pt, mc, have_box, occ_lo, occ_hi, tid);
    if (tid < TM && sl.active) {
        const float sigma = s.sig[tid];
        const float alpha = 1.0f - expf(-sigma * sl.dt);
        const float T = 1.0f - sl.ws;
        const float w = alpha * T;
        sl.ws += w;
        sl.depth += w * sl.t;  // sl.t is already the post-sample t (deltas[1])
        sl.r += w * s.rgb[tid];
        sl.gch += w * s.rgb[TM + tid];
        sl.b += w * s.rgb[2 * TM + tid];
        sl.nsamp += 1;
        if (a.valid_samples) warp_agg_add(a.valid_samples, sl.frame, 1);
        int D = 0;  // death index (1-based sample position), 0 = still alive
        bool suspend = false;
        if (T < a.T_thresh) D = sl.nsamp;
        else if (sl.nsamp >= sl.cap) suspend = true;
        else if (!march_next_nl(mc, sl.g, sl.far_m, sl.t, sl.px, sl.py, sl.pz, sl.dt)) D = sl.nsamp + 1;
        if (D) {
            finalize_ray(a, sl, true);
            if (a.pass == 1) warp_agg_add(a.hist, sl.frame * (a.max_steps + 2) + D, 1);
            sl.active = false;
        } else if (suspend) {
            if (a.pass == 1) {
                finalize_ray(a, sl, false);  // raw depth: pass 2 keeps accumulating
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

// copy the coarse occupancy words into shared memory and point the marcher at them
template <class SmemT>
__device__ __forceinline__ void install_coarse(const HeadArgs &a, SmemT &s, MarchConst &mc, int tid, int nthreads) {
    if (a.coarse_bits && a.coarse_words <= HEAD_COARSE_WORDS) {
        for (int i = tid; i < a.coarse_words; i += nthreads) s.coarse[i] = a.coarse_bits[i];
        mc.coarse = s.coarse;
    } else {
        mc.coarse = a.coarse_bits;   // too large for the smem slot (cascade > 2): read through L1/L2
    }
}

}  // namespace headc
}  // namespace gfpp
