// Multi-subunit complexes for K2+K3 (north-star item 3; EXTENSION - the reference has no complex handling and
// expects one LR row per subunit pair, SURVEY.md "gaps").  A complex occupies a "virtual" column of the warp's
// shared D slab; after phase 1 its value is aggregated from the subunit columns,
//   rule 0  min over subunits              (CellPhoneDB convention)
//   rule 1  geometric mean of max(v, 0)    (CellChat convention)
// and it counts as expressed in the cell only if every subunit is.  Phase 2 then scores pairs that name the
// virtual column exactly like any other column.
#pragma once
#include <float.h>
#include <stdint.h>

namespace laris {

struct ComplexTable {
    const int32_t* ptr = nullptr;    // [nc+1] offsets into cols
    const int32_t* cols = nullptr;   // subunit columns
    const int32_t* vcol = nullptr;   // [nc] virtual column of each complex
    int nc = 0;
    int rule = 0;
};

// Dw: the cell's D slab; Bw: its expression bitmap (SIGNED) - otherwise the flag rides in the sign bit of D >= 0.
// Caller brackets the call with __syncwarp().
template <bool SIGNED>
__device__ __forceinline__ void complex_pass(const ComplexTable& cx, float* Dw, uint32_t* Bw, int lane) {
    for (int c = lane; c < cx.nc; c += 32) {
        const int s = cx.ptr[c], e = cx.ptr[c + 1];
        float mn = FLT_MAX, lsum = 0.f;
        bool all = e > s, zero = false;
        for (int t = s; t < e; ++t) {
            const int col = cx.cols[t];
            float v = Dw[col];
            bool f;
            if (SIGNED) {
                f = (Bw[col >> 5] >> (col & 31)) & 1u;
            } else {
                f = (__float_as_uint(v) >> 31) != 0u;
                v = fabsf(v);
            }
            all &= f;
            mn = fminf(mn, v);
            if (v > 0.f) lsum += logf(v); else zero = true;
        }
        float val = e > s ? (cx.rule == 0 ? mn : (zero ? 0.f : expf(lsum / (float)(e - s)))) : 0.f;
        const int vc = cx.vcol[c];
        if (SIGNED) {
            Dw[vc] = val;
            if (all) atomicOr(&Bw[vc >> 5], 1u << (vc & 31));
        } else {
            Dw[vc] = all ? -val : val;
        }
    }
}

}  // namespace laris
