// One node of the BVH2 -> BVH8 collapse: which (up to 8) descendants of a binary-tree node become its children, which slot
// each one takes, and the 8-bit quantised child boxes.  __host__ __device__ and free of library calls whose results could
// differ between glibc and the CUDA math library, so the host flattener (dt_flatten.cu) and the GPU flattener
// (dt_flatten_gpu.cu) emit the same bytes.
#pragma once
#include <float.h>
#include <math.h>
#include <string.h>
#include "dt_flatten.h"

#define DT_COLLAPSE_OK 0
#define DT_COLLAPSE_ERR_EXPONENT 1
#define DT_COLLAPSE_ERR_LEAF 2

__host__ __device__ inline float dt_box_area(const float* mn, const float* mx) {
    float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    if (!(dx >= 0) || !(dy >= 0) || !(dz >= 0)) return 0.f;
    return dx * dy + dy * dz + dz * dx;
}
__host__ __device__ inline bool dt_b2_is_leaf(const DtB2Node* b2, int n) { return b2[n].left < 0 || b2[n].count <= 1; }   // every leaf child is ONE primitive

// ceil(log2(x)) for finite x > 0 without calling log2: x = f * 2^e with f in [0.5, 1)
__host__ __device__ inline int dt_ceil_log2(double x) { int e; const double f = frexp(x, &e); return f == 0.5 ? e - 1 : e; }
__host__ __device__ inline float dt_next_below(float x) {                    // nextafterf(x, -FLT_MAX) for finite x
    if (x == 0.0f) return -1.401298464e-45f;
    uint32_t b; memcpy(&b, &x, 4);
    b = (b & 0x80000000u) ? b + 1u : b - 1u;
    float r; memcpy(&r, &b, 4); return r;
}

// node: everything except child_base / prim_base (imask / lmask say which slots hold inner / leaf children);
// child_in_slot[s] = binary-tree node in slot s or -1.
__host__ __device__ inline int dt_collapse_node(const DtB2Node* b2, int b2node, DtNode8& node, int child_in_slot[8]) {
    const DtB2Node& root = b2[b2node];
    int ch[8]; int n_ch = 0;
    if (dt_b2_is_leaf(b2, b2node)) ch[n_ch++] = b2node;            // tiny tree: the root itself is the only (leaf) child
    else { ch[n_ch++] = root.left; ch[n_ch++] = root.right; }
    while (n_ch < 8) {                                             // open the child with the largest surface area
        int best = -1; float best_area = -1.f;
        for (int k = 0; k < n_ch; k++) {
            if (dt_b2_is_leaf(b2, ch[k])) continue;
            const float a = dt_box_area(b2[ch[k]].mn, b2[ch[k]].mx);
            if (a > best_area) { best_area = a; best = k; }
        }
        if (best < 0) break;
        const int n = ch[best];
        ch[best] = b2[n].left;
        ch[n_ch++] = b2[n].right;
    }
    // node bounds = union of children
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = 0; k < n_ch; k++) for (int a = 0; a < 3; a++) {
        if (b2[ch[k]].mn[a] < mn[a]) mn[a] = b2[ch[k]].mn[a];
        if (b2[ch[k]].mx[a] > mx[a]) mx[a] = b2[ch[k]].mx[a];
    }
    for (int a = 0; a < 3; a++) if (!(mn[a] <= mx[a])) { mn[a] = 0.f; mx[a] = 0.f; }   // NaN / empty guard
    // slot assignment: slot s prefers the child towards corner (s&4 ? +x : -x, s&2 ? +y : -y, s&1 ? +z : -z)
    const float cen[3] = {0.5f * (mn[0] + mx[0]), 0.5f * (mn[1] + mx[1]), 0.5f * (mn[2] + mx[2])};
    int slot_of[8];
    for (int k = 0; k < 8; k++) { slot_of[k] = -1; child_in_slot[k] = -1; }
    float cost[8][8];
    for (int k = 0; k < n_ch; k++) {
        const DtB2Node& c = b2[ch[k]];
        const float d[3] = {0.5f * (c.mn[0] + c.mx[0]) - cen[0], 0.5f * (c.mn[1] + c.mx[1]) - cen[1], 0.5f * (c.mn[2] + c.mx[2]) - cen[2]};
        for (int s = 0; s < 8; s++) cost[k][s] = ((s & 4) ? d[0] : -d[0]) + ((s & 2) ? d[1] : -d[1]) + ((s & 1) ? d[2] : -d[2]);
    }
    int slot_child[8];                                             // index into ch per slot
    for (int s = 0; s < 8; s++) slot_child[s] = -1;
    for (int it = 0; it < n_ch; it++) {
        float bc = -FLT_MAX; int bk = -1, bs = -1;
        for (int k = 0; k < n_ch; k++) {
            if (slot_of[k] >= 0) continue;
            for (int s = 0; s < 8; s++) {
                if (slot_child[s] >= 0) continue;
                float cst = cost[k][s];
                if (!(cst == cst)) cst = 0.f;
                if (cst > bc || bk < 0) { bc = cst; bk = k; bs = s; }
            }
        }
        slot_of[bk] = bs; slot_child[bs] = bk;
    }
    // quantisation frame.  Child boxes are rounded OUTWARD with a 1/128-step margin on both sides: the kernel's node
    // test evaluates plane distances with an absolute error of up to 2^-9 quantisation steps (dt_traverse.cuh,
    // dt_byte_m).  The frame origin sits a margin below the node minimum so the margin also holds at q = 0.
    memset(&node, 0, sizeof node);
    const double margin = 1.0 / 128.0;
    int e[3];
    for (int a = 0; a < 3; a++) {
        const double ext = (double)mx[a] - (double)mn[a];
        int ea = -120;
        if (ext > 0) ea = dt_ceil_log2(ext / 254.0);
        if (ea < -120) ea = -120;
        if (ea > 100) ea = 100;
        e[a] = ea;
    }
    uint8_t qlo[3][8], qhi[3][8];
    float origin[3];
    for (int a = 0; a < 3; a++) {
        for (;;) {
            bool ok = true;
            const double sc = ldexp(1.0, e[a]);
            float pf = (float)((double)mn[a] - 1.5 * margin * sc);
            if ((double)pf > (double)mn[a] - margin * sc) pf = dt_next_below(pf);
            if (!(pf == pf) || !(fabsf(pf) <= FLT_MAX)) pf = mn[a];
            const double p = (double)pf;
            for (int s = 0; s < 8 && ok; s++) {
                const int k = slot_child[s];
                if (k < 0) { qlo[a][s] = 0; qhi[a][s] = 0; continue; }
                double lo = b2[ch[k]].mn[a], hi = b2[ch[k]].mx[a];
                if (!(lo <= hi)) { lo = p; hi = p; }
                double ql = floor((lo - p) / sc - margin), qh2 = ceil((hi - p) / sc + margin);
                if (ql < 0) ql = 0;
                while (p + (ql + margin) * sc > lo && ql > 0) ql -= 1;
                while (p + (qh2 - margin) * sc < hi) qh2 += 1;
                if (qh2 > 255 || ql > 255) { ok = false; break; }
                qlo[a][s] = (uint8_t)ql; qhi[a][s] = (uint8_t)qh2;
            }
            if (ok) { origin[a] = pf; break; }
            e[a]++;
            if (e[a] > 100) return DT_COLLAPSE_ERR_EXPONENT;
        }
    }
    node.px = origin[0]; node.py = origin[1]; node.pz = origin[2];
    node.ex = (uint8_t)(e[0] + 127); node.ey = (uint8_t)(e[1] + 127); node.ez = (uint8_t)(e[2] + 127);
    for (int s = 0; s < 8; s++) {
        const int k = slot_child[s];
        if (k < 0) continue;
        const int c = ch[k];
        child_in_slot[s] = c;
        if (dt_b2_is_leaf(b2, c)) {
            if (b2[c].count != 1) return DT_COLLAPSE_ERR_LEAF;
            node.lmask |= (uint8_t)(1u << s);
        } else node.imask |= (uint8_t)(1u << s);
        node.qlox[s] = qlo[0][s]; node.qloy[s] = qlo[1][s]; node.qloz[s] = qlo[2][s];
        node.qhix[s] = qhi[0][s]; node.qhiy[s] = qhi[1][s]; node.qhiz[s] = qhi[2][s];
    }
    return DT_COLLAPSE_OK;
}
