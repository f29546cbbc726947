// Per-component MNA stamp: the arithmetic of nodal/models.py:13-214 for ONE component,
// written once as a __host__ __device__ function so the GPU kernel (stamp.cu) and the
// CPU-side unit check (tests/host_check/stamp_host.cu) share it.
//
// Output: up to NODAL_MAX_TRIPLES (row, col, val) entries.  Right-hand-side
// contributions carry col == n.  The reference mixes '=' and '+=' writes inside one
// component (e.g. write_VCVS: incidence '=' then control '+='); those only ever collide
// inside the component's own branch row, so they are resolved here in a 4-entry ordered
// map and emitted as final values.  Entries of different components that share a slot
// (conductance sums in the KCL block) are all '+=' and are summed in component order by
// the CSR build.
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

#define NODAL_MAX_TRIPLES 6

struct StampOut {
    int32_t row[NODAL_MAX_TRIPLES];
    int32_t col[NODAL_MAX_TRIPLES];
    double val[NODAL_MAX_TRIPLES];
    int count;
};

struct BranchRow {  // ordered (col -> val) map of the component's own branch row
    int32_t col[4];
    double val[4];
    int cnt;
    __host__ __device__ void assign(int32_t c, double v) {  // G[r, c] = v
        for (int i = 0; i < cnt; ++i)
            if (col[i] == c) { val[i] = v; return; }
        col[cnt] = c; val[cnt] = v; ++cnt;
    }
    __host__ __device__ void accumulate(int32_t c, double v) {  // G[r, c] += v
        for (int i = 0; i < cnt; ++i)
            if (col[i] == c) { val[i] = val[i] + v; return; }
        col[cnt] = c; val[cnt] = 0.0 + v; ++cnt;
    }
};

__host__ __device__ inline void emit(StampOut& o, int32_t r, int32_t c, double v) {
    o.row[o.count] = r; o.col[o.count] = c; o.val[o.count] = v; ++o.count;
}

// type codes: NODAL_T_* of include/nodal_b200.h.  a,b,c,d: node rows or -1 (ground).
// drv_value: resistance of the driving resistor (CCVS/CCCS only).
__host__ __device__ inline void stamp_component(int type, double value, int32_t a, int32_t b,
                                                int32_t c, int32_t d, double drv_value,
                                                int32_t branch, int32_t kcl, int32_t n,
                                                StampOut& o) {
    o.count = 0;
    if (type == 0) {  // R -- models.py:13-24
        const double g = 1.0 / value;
        if (a >= 0) emit(o, a, a, g);
        if (b >= 0) emit(o, b, b, g);
        if (a >= 0 && b >= 0) { emit(o, a, b, -g); emit(o, b, a, -g); }
        return;
    }
    if (type == 1) {  // A -- models.py:27-32 (rhs lives in column n)
        if (a >= 0) emit(o, a, n, value);
        if (b >= 0) emit(o, b, n, -value);
        return;
    }
    const int32_t r = kcl + branch;  // branch row == branch column (models.py:38)
    // KCL-row incidence: G[a, r] = -1 ; G[b, r] = +1  (the later write wins if a == b)
    if (a >= 0 && a != b) emit(o, a, r, -1.0);
    if (b >= 0) emit(o, b, r, 1.0);
    BranchRow br; br.cnt = 0;
    if (type == 6) {  // CCCS -- models.py:161-199
        br.assign(r, 1.0);
        if (c >= 0) br.assign(c, value / drv_value);
        if (d >= 0) br.assign(d, -value / drv_value);
    } else {
        if (type == 2) emit(o, r, n, value);  // E: A[r] += V -- models.py:39
        if (a >= 0) br.assign(a, 1.0);        // G[r, a] = 1
        if (b >= 0) br.assign(b, -1.0);       // G[r, b] = -1
        if (type == 3 || type == 4) {         // VCVS and VCCS (nodal.py:377-380) -- models.py:71-78
            if (c >= 0) br.accumulate(c, -value);
            if (d >= 0) br.accumulate(d, value);
        } else if (type == 5) {               // CCVS -- models.py:140-145 (assignment)
            if (c >= 0) br.assign(c, value / drv_value);
            if (d >= 0) br.assign(d, -value / drv_value);
        }
    }
    for (int i = 0; i < br.cnt; ++i) emit(o, r, br.col[i], br.val[i]);
}

// Largest number of entries a type can emit (host side uses this to pick the stride).
__host__ __device__ inline int stamp_max_entries(int type) {
    switch (type) {
        case 0: return 4;
        case 1: return 2;
        case 2: return 5;
        case 6: return 5;
        default: return 6;
    }
}
