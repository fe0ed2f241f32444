// Device helpers shared by every kernel.  The whole library is compiled with -fmad=false: the reference is an
// x86-64 build without FMA contraction, so every a*b+c below rounds twice, as there.
#pragma once
#include <cuda_runtime.h>
#include "goicp_dev.h"

#define GOICP_FULL 0xFFFFFFFFu

// ROUND((v-min)*scale), jly_3ddt.cpp:30,1143-1145: double arithmetic, truncation toward zero.
__device__ __forceinline__ int vox_round(float v, double mn, double scale) {
    return (int)(((double)v - mn) * scale + 0.5);
}
__device__ __forceinline__ int vox_round_d(double v, double mn, double scale) {
    return (int)((v - mn) * scale + 0.5);
}

// DT3D::Distance (jly_3ddt.cpp:1139-1191) for a position already rounded to float (every caller passes floats
// widened to double: jly_goicp.cpp:123-135,368-369,601-604).
__device__ __forceinline__ float dt_distance(const GridDev& g, const float* __restrict__ dist, float fx, float fy, float fz) {
    const int S = g.S;
    int x = vox_round(fx, g.xMin, g.scale), y = vox_round(fy, g.yMin, g.scale), z = vox_round(fz, g.zMin, g.scale);
    if ((unsigned)x < (unsigned)S && (unsigned)y < (unsigned)S && (unsigned)z < (unsigned)S)
        return __ldg(dist + ((z * S + y) * S + x));
    float a = 0.f, b = 0.f, c = 0.f;
    if (x < 0) { a = (float)x; x = 0; } else if (x >= S) { a = (float)(x - S + 1); x = S - 1; }
    if (y < 0) { b = (float)y; y = 0; } else if (y >= S) { b = (float)(y - S + 1); y = S - 1; }
    if (z < 0) { c = (float)z; z = 0; } else if (z >= S) { c = (float)(z - S + 1); z = S - 1; }
    return (float)((double)sqrtf(a * a + b * b + c * c) / g.scale + (double)__ldg(dist + ((z * S + y) * S + x)));
}
__device__ __forceinline__ float dt_distance_d(const GridDev& g, double dx, double dy, double dz, int* ox, int* oy, int* oz) {
    const int S = g.S;
    int x = vox_round_d(dx, g.xMin, g.scale), y = vox_round_d(dy, g.yMin, g.scale), z = vox_round_d(dz, g.zMin, g.scale);
    *ox = x; *oy = y; *oz = z;
    if ((unsigned)x < (unsigned)S && (unsigned)y < (unsigned)S && (unsigned)z < (unsigned)S)
        return g.dist[(z * S + y) * S + x];
    float a = 0.f, b = 0.f, c = 0.f;
    if (x < 0) { a = (float)x; x = 0; } else if (x >= S) { a = (float)(x - S + 1); x = S - 1; }
    if (y < 0) { b = (float)y; y = 0; } else if (y >= S) { b = (float)(y - S + 1); y = S - 1; }
    if (z < 0) { c = (float)z; z = 0; } else if (z >= S) { c = (float)(z - S + 1); z = S - 1; }
    return (float)((double)sqrtf(a * a + b * b + c * c) / g.scale + (double)g.dist[(z * S + y) * S + x]);
}

// register-argument forms of the two lookups for the hot loop (grid constants hoisted out of the PairDev)
template <bool LDG = true>
__device__ __forceinline__ float dt_distance_v(int S, double x0, double y0, double z0, double scale, const float* dist, float fx, float fy, float fz) {
    int x = vox_round(fx, x0, scale), y = vox_round(fy, y0, scale), z = vox_round(fz, z0, scale);
    if ((unsigned)x < (unsigned)S && (unsigned)y < (unsigned)S && (unsigned)z < (unsigned)S)
        return LDG ? __ldg(dist + ((z * S + y) * S + x)) : dist[(z * S + y) * S + x];
    float a = 0.f, b = 0.f, c = 0.f;
    if (x < 0) { a = (float)x; x = 0; } else if (x >= S) { a = (float)(x - S + 1); x = S - 1; }
    if (y < 0) { b = (float)y; y = 0; } else if (y >= S) { b = (float)(y - S + 1); y = S - 1; }
    if (z < 0) { c = (float)z; z = 0; } else if (z >= S) { c = (float)(z - S + 1); z = S - 1; }
    return (float)((double)sqrtf(a * a + b * b + c * c) / scale + (double)(LDG ? __ldg(dist + ((z * S + y) * S + x)) : dist[(z * S + y) * S + x]));
}
__device__ __forceinline__ int clamp_cell_v(int S, double x0, double y0, double z0, double scale, const int* __restrict__ vcell, float fx, float fy, float fz) {
    int x = vox_round(fx, x0, scale), y = vox_round(fy, y0, scale), z = vox_round(fz, z0, scale);
    x = min(max(x, 0), S - 1); y = min(max(y, 0), S - 1); z = min(max(z, 0), S - 1);
    return __ldg(vcell + ((z * S + y) * S + x));
}

__device__ __forceinline__ int clamp_vox_v(int S, double x0, double y0, double z0, double scale, float fx, float fy, float fz) {
    int x = vox_round(fx, x0, scale), y = vox_round(fy, y0, scale), z = vox_round(fz, z0, scale);
    x = min(max(x, 0), S - 1); y = min(max(y, 0), S - 1); z = min(max(z, 0), S - 1);
    return (z * S + y) * S + x;
}

// checkCompatibility's voxel (jly_goicp.cpp:976-984): same rounding, clamped INTO the grid; returns the compact id
// of the closest occupied cell (emptyCells), ncells if that voxel is unresolved.
__device__ __forceinline__ int clamp_cell(const GridDev& g, float fx, float fy, float fz) {
    const int S = g.S;
    int x = vox_round(fx, g.xMin, g.scale), y = vox_round(fy, g.yMin, g.scale), z = vox_round(fz, g.zMin, g.scale);
    x = min(max(x, 0), S - 1); y = min(max(y, 0), S - 1); z = min(max(z, 0), S - 1);
    return __ldg(g.vcell + ((z * S + y) * S + x));
}

// FP32 fast path of the voxel index (GridDev.vf*): kx,ky,kz = bit patterns of fma(p, vfScale, C) with
// C = (float)((trans - min)*scale + vfMagic).  Returns the linear voxel index, or -1 when any axis lies in the
// ambiguity zone of a rounding boundary or outside the grid (the caller then runs the exact FP64 form).
struct VoxFast { float sc; int sh; unsigned bias, mask, zone; };
__device__ __forceinline__ VoxFast vox_fast_of(const GridDev& g) { VoxFast v; v.sc = g.vfScale; v.sh = g.vfShift; v.bias = g.vfBias; v.mask = g.vfMask; v.zone = g.vfZone; return v; }
__device__ __forceinline__ float vox_fast_c(double magic, float trans, double mn, double scale) { return (float)(((double)trans - mn) * scale + magic); }
__device__ __forceinline__ int vox_fast(const VoxFast& vf, int S, float px, float py, float pz, float Cx, float Cy, float Cz) {
    const unsigned kx = __float_as_uint(__fmaf_rn(px, vf.sc, Cx)), ky = __float_as_uint(__fmaf_rn(py, vf.sc, Cy)), kz = __float_as_uint(__fmaf_rn(pz, vf.sc, Cz));
    const unsigned x = (kx >> vf.sh) - vf.bias, y = (ky >> vf.sh) - vf.bias, z = (kz >> vf.sh) - vf.bias;
    const unsigned fr = min(min(kx & vf.mask, ky & vf.mask), kz & vf.mask);
    const unsigned mx = max(max(x, y), z);
    if (fr < vf.zone || mx >= (unsigned)S) return -1;
    return (int)((z * S + y) * S + x);
}

// nearestNeighbor (jly_goicp.cpp:1200-1211) inside compact cell `cell` + one term of compareNeighbors(false, ...) (:1250-1288):
// |neighbours of data point i - neighbours of the closest model point of the cell| (model point 0 if the cell is empty).
__device__ __forceinline__ int nb_diff(const PairDev& P, int cell, int i, float ax, float ay, float az) {
    double minD = 100.0; int ind = 0;
    for (int k = cell < P.g.ncells ? __ldg(P.cell_start + cell) : 0, e = cell < P.g.ncells ? __ldg(P.cell_start + cell + 1) : 0; k < e; ++k) {
        const int p = __ldg(P.cell_pts + k);
        const double a = (double)(ax - __ldg(P.mx + p)), b = (double)(ay - __ldg(P.my + p)), c = (double)(az - __ldg(P.mz + p));
        const double d = sqrt(a * a + b * b + c * c);
        if (d < minD) { ind = p; minD = d; }
    }
    return abs(__ldg(P.nbD + i) - __ldg(P.nbM + ind));
}

// Second tier: a position up to GOICP_OVLIM voxels outside the grid.  Returns the clamped linear index and the squared
// voxel overshoot a^2+b^2+c^2 of DT3D::Distance (jly_3ddt.cpp:1150-1190); false when the position is in the ambiguity zone or
// further out (the caller then runs the exact FP64 form).  GridDev.ovl[s] = (double)sqrtf(s) / scale.
__device__ __forceinline__ bool vox_near(const VoxFast& vf, int S, float px, float py, float pz, float Cx, float Cy, float Cz, int* idx, int* s2) {
    const unsigned kx = __float_as_uint(__fmaf_rn(px, vf.sc, Cx)), ky = __float_as_uint(__fmaf_rn(py, vf.sc, Cy)), kz = __float_as_uint(__fmaf_rn(pz, vf.sc, Cz));
    int xi = (int)(kx >> vf.sh) - (int)vf.bias, yi = (int)(ky >> vf.sh) - (int)vf.bias, zi = (int)(kz >> vf.sh) - (int)vf.bias;
    const unsigned fr = min(min(kx & vf.mask, ky & vf.mask), kz & vf.mask);
    // the mantissa holds floor(t + 0.5); ROUND (jly_3ddt.cpp:30) truncates toward zero, so below the grid's low edge the
    // reference's voxel is one higher (outside the ambiguity zone t + 0.5 is never an integer): (-1, 0) -> 0, (-2, -1) -> -1 ...
    xi += (xi < 0); yi += (yi < 0); zi += (zi < 0);
    const int cx = min(max(xi, 0), S - 1), cy = min(max(yi, 0), S - 1), cz = min(max(zi, 0), S - 1);
    const int ax = xi - cx, ay = yi - cy, az = zi - cz;
    *s2 = ax * ax + ay * ay + az * az;
    *idx = (cz * S + cy) * S + cx;
    return fr >= vf.zone && max(max(ax, ay), az) <= GOICP_OVLIM && min(min(ax, ay), az) >= -GOICP_OVLIM;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(GOICP_FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(GOICP_FULL, v); }

// k-th smallest (1-based) of n non-negative floats v[0..n) by a warp: bitwise radix select on the float bit patterns
// (replaces intro_select, jly_sorting.hpp:229-313).  Returns the bit pattern T of the k-th smallest and in *need_eq
// how many elements equal to T belong to the k smallest (they are taken in index order).
__device__ __forceinline__ unsigned warp_select_kth(const float* v, int n, int k, int lane, int* need_eq) {
    unsigned prefix = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const unsigned hi = (bit == 31) ? 0u : (0xFFFFFFFFu << (bit + 1));
        int c0 = 0;
        for (int i = lane; i < n; i += 32) {
            unsigned u = __float_as_uint(v[i]);
            c0 += ((u & hi) == prefix && !((u >> bit) & 1u)) ? 1 : 0;
        }
        c0 = warp_sum_i(c0);
        if (k > c0) { k -= c0; prefix |= (1u << bit); }
    }
    *need_eq = k;
    return prefix;
}

// Sum over the k smallest of v[0..n) (set chosen by warp_select_kth) of f(v) for the two bound sums of
// InnerBnB (jly_goicp.cpp:393-415).  Tree order: trimmed sums are compared at tolerance (the reference adds them in
// intro_select's permutation order, which is not reproducible in closed form).
// mask (optional): mask[i] = 1 when point i is in the k-smallest set -- the per-cube point-inclusion mask (goicp_eval_inclusion).
__device__ __forceinline__ void warp_trimmed_sums(const float* v, int n, int k, int lane, int norm, float mtd, float* ub, float* lb, uint8_t* mask = nullptr) {
    int need_eq;
    const unsigned T = warp_select_kth(v, n, k, lane, &need_eq);
    float su = 0.f, sl = 0.f;
    int eq_seen = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const float m = (i < n) ? v[i] : 0.f;
        const unsigned u = __float_as_uint(m);
        const bool valid = i < n;
        const unsigned eqm = __ballot_sync(GOICP_FULL, valid && u == T);
        const int rank = eq_seen + __popc(eqm & ((1u << lane) - 1u));
        const bool inc = valid && (u < T || (u == T && rank < need_eq));
        eq_seen += __popc(eqm);
        if (mask && valid) mask[i] = inc ? 1 : 0;
        if (inc) {
            su += (norm == 2) ? m * m : m;
            const float d = m - mtd;
            if (d > 0.f) sl += (norm == 2) ? d * d : d;
        }
    }
    *ub = warp_sum(su);
    *lb = warp_sum(sl);
}
