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
__device__ __forceinline__ bool may_hit_occupied(bool have_box, const float (&occ_lo)[3], const float (&occ_hi)[3],
                                                 const RayGeom &g, float near, float far) {
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


}  // namespace headc
}  // namespace gfpp
