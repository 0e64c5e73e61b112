// T6 - rows of the reference's long result table, ordered and gathered on the device.
//
// The reference melts the [pair][receiver][sender] score cube into a long frame, sorts it by the pre-weight score
// (laris/tools/_utils.py:1422) and then by the final score (:1472), both descending.  The host mirror of that
// (two stable numpy argsorts over the rows with a positive score, zero rows after them in melt order) cost ~50 ms per
// run at 2000 pairs x 40 x 40 types, plus ~25 ms of fancy-index gathers for the numeric columns.  Here: two stable
// LSD radix sorts (cub::DeviceRadixSort, CUDA toolkit) of all Pt*L rows by explicit 64-bit keys, then ONE gather pass
// that emits the numeric columns and the per-row pair / sender / receiver codes in table order.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace laris {

// ascending order of this key == descending order of x; -0.0 and +0.0 compare equal (as in numpy)
__device__ __forceinline__ uint64_t descending_key(double x) {
    if (x == 0.0) x = 0.0;
    uint64_t b = (uint64_t)__double_as_longlong(x);
    b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    return ~b;
}

// melt position m = (pair, receiver, sender) -> element f = (pair, sender, receiver) of the [Pt][C][C] tables.
// key1 = pre-weight score of the rows with a positive final score; every other row gets the largest key, so the first
// sort leaves them behind the positive rows in melt order.
__global__ void result_order_keys_kernel(const double* __restrict__ inter, const double* __restrict__ flat, int64_t total,
                                         int C, uint64_t* __restrict__ key1, int32_t* __restrict__ elem,
                                         int32_t* __restrict__ exotic) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= total) return;
    const int64_t L = (int64_t)C * C;
    const int64_t pair = m / L;
    const int rem = (int)(m - pair * L);
    const int r = rem / C, s = rem - r * C;
    const int64_t f = pair * L + (int64_t)s * C + r;
    const double fl = flat[f];
    const double in = inter[f];
    if (!(fl >= 0.0) || (fl > 0.0 && in != in)) atomicOr(exotic, 1);       // negative / NaN scores: the host orders them
    key1[m] = fl > 0.0 ? descending_key(in) : ~0ull;
    elem[m] = (int32_t)f;
}

__global__ void result_order_key2_kernel(const double* __restrict__ flat, const int32_t* __restrict__ elem, int64_t total,
                                         uint64_t* __restrict__ key2) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < total) key2[j] = descending_key(flat[elem[j]]);
}

// one pass over the ordered rows: numeric columns (null sources are skipped) and the codes the string columns are
// decoded from
__global__ void result_rows_gather_kernel(const int32_t* __restrict__ order, int64_t total, int C,
                                          const double* __restrict__ flat, const double* __restrict__ p_rows,
                                          const double* __restrict__ fdr_rows, const double* __restrict__ nlog_rows,
                                          const double* __restrict__ extra, double* __restrict__ score_o,
                                          double* __restrict__ p_o, double* __restrict__ fdr_o, double* __restrict__ nlog_o,
                                          double* __restrict__ extra_o, int32_t* __restrict__ pair_o,
                                          int32_t* __restrict__ send_o, int32_t* __restrict__ recv_o) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= total) return;
    const int32_t f = order[j];
    const int L = C * C;
    const int pair = f / L, rem = f - pair * L;
    const int s = rem / C;
    pair_o[j] = pair;
    send_o[j] = s;
    recv_o[j] = rem - s * C;
    score_o[j] = flat[f];
    if (p_rows) p_o[j] = p_rows[f];
    if (fdr_rows) fdr_o[j] = fdr_rows[f];
    if (nlog_rows) nlog_o[j] = nlog_rows[f];
    if (extra) extra_o[j] = extra[f];
}

}  // namespace laris

using namespace laris;

extern "C" int laris_result_row_order(const double* inter, const double* flat, int32_t Pt, int32_t C, int32_t* order,
                                      int32_t* exotic, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LARIS_CHECK_ARG(inter && flat && order && exotic && Pt >= 1 && C >= 1, "laris_result_row_order: bad arguments");
    const int64_t total = (int64_t)Pt * C * C;
    LARIS_CHECK_ARG(total < (int64_t)0x7fffffff, "laris_result_row_order: %lld rows exceed the 2^31 - 1 limit",
                    (long long)total);
    Scratch<uint64_t> key_a(st), key_b(st);
    Scratch<int32_t> elem_a(st), elem_b(st);
    Scratch<uint8_t> temp(st);
    LARIS_CUDA(key_a.alloc((size_t)total));
    LARIS_CUDA(key_b.alloc((size_t)total));
    LARIS_CUDA(elem_a.alloc((size_t)total));
    LARIS_CUDA(elem_b.alloc((size_t)total));
    size_t temp_bytes = 0;
    LARIS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, key_a.p, key_b.p, elem_a.p, elem_b.p, (int)total, 0, 64,
                                               st));
    LARIS_CUDA(temp.alloc(temp_bytes));
    LARIS_CUDA(cudaMemsetAsync(exotic, 0, sizeof(int32_t), st));
    const unsigned grid = (unsigned)ceil_div(total, 256);
    result_order_keys_kernel<<<grid, 256, 0, st>>>(inter, flat, total, C, key_a.p, elem_a.p, exotic);
    LARIS_LAUNCH_CHECK();
    LARIS_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, key_a.p, key_b.p, elem_a.p, elem_b.p, (int)total, 0, 64,
                                               st));
    note_launch();
    result_order_key2_kernel<<<grid, 256, 0, st>>>(flat, elem_b.p, total, key_a.p);
    LARIS_LAUNCH_CHECK();
    LARIS_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, key_a.p, key_b.p, elem_b.p, order, (int)total, 0, 64, st));
    note_launch();
    return LARIS_OK;
}

extern "C" int laris_result_rows_gather(const int32_t* order, int64_t total, int32_t C, const double* flat,
                                        const double* p_rows, const double* fdr_rows, const double* nlog_rows,
                                        const double* extra, double* score_o, double* p_o, double* fdr_o, double* nlog_o,
                                        double* extra_o, int32_t* pair_o, int32_t* send_o, int32_t* recv_o, void* stream) {
    LARIS_CHECK_ARG(order && flat && score_o && pair_o && send_o && recv_o && total >= 0 && C >= 1,
                    "laris_result_rows_gather: bad arguments");
    LARIS_CHECK_ARG((!p_rows || p_o) && (!fdr_rows || fdr_o) && (!nlog_rows || nlog_o) && (!extra || extra_o),
                    "laris_result_rows_gather: a source column without its output");
    if (total == 0) return LARIS_OK;
    result_rows_gather_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
        order, total, C, flat, p_rows, fdr_rows, nlog_rows, extra, score_o, p_o, fdr_o, nlog_o, extra_o, pair_o, send_o,
        recv_o);
    LARIS_LAUNCH_CHECK();
    return LARIS_OK;
}
