// tcgen05 variant of the cell-type-pair contraction (placeholder until the
// tensor-core kernel lands: reports an error instead of silently falling back).
#include "common.cuh"

namespace laris {
int celltype_pair_contract_mma(const float*, int64_t, int64_t, int32_t, const float*, int32_t, double*, cudaStream_t) {
    set_error("laris_celltype_pair_contract: impl=0 (tcgen05) is not built in this revision; pass impl=1");
    return LARIS_E_LIMIT;
}
}  // namespace laris
