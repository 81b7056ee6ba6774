// ee_sym_types.h -- plain structs shared by the pair-symmetric kernels (ee_sym.cuh) and the host engine (ee_engine.h).
#pragma once
#include <cstdint>

namespace ee {

struct SymItem {  // one work item: chunks [c0, c0 + nc) of tile row ti; the i-side partial sum goes to slot `slot`
    int ti, c0, nc, slot;
};

// What the reduce kernel needs to know about a rank's share (canonical unit space, see ee_sym.cuh).
struct SymShare {
    long long u_lo, u_hi;  // units owned by this rank
    int tile;              // bodies per I-tile
    int nt;                // tile rows
    int nch;               // chunks on the j axis (n / 32)
};

// first canonical unit of tile row ti: rows r = 0..ti-1 hold nch - r*cpt units each (cpt = chunks per tile)
#ifdef __CUDACC__
__host__ __device__
#endif
inline long long sym_row_unit(long long ti, long long nch, long long cpt) {
    return ti * nch - cpt * (ti * (ti - 1) / 2);
}

}  // namespace ee
