// moira_groups.cu -- the bookkeeping half of --collapse on the device (moira/moira.py:459-475, 491-504):
// labels (equal label <=> equal sequence, from moira_dedup.cu) + expected errors -> the reference's groups in order of
// first appearance, the representative of every group (first read with the STRICTLY smallest ee, :466), the order of
// names (every new representative goes to the front, every other read to the back, :470-475) and the groups by
// abundance, largest first, ties in order of first appearance (:492).  The host version (moira_collapse_labels) walks the
// reads three times in input order with random accesses -- 0.3 s for 5 M reads, eight times the filter kernels of the
// same reads; here it is a handful of linear passes:
//
//   first[label]   = smallest read index of the label            (atomicMin)
//   group id       = exclusive sum of "this read is its label's first"  -> first-appearance numbering
//   size[g]        = atomic count;  member_start = exclusive sum
//   (g, r) stable radix sort by g                                 -> every group's reads side by side, in input order
//   running minimum of ee inside each group (scan by key)         -> "record breakers" = the successive representatives
//   breakers go to the front in reverse order, the others behind in input order; the last breaker is the representative
//   groups by ~size, stable radix sort                            -> abundance order
//
// The sorts and scans are CUB's device-wide primitives (part of the CUDA toolkit, like cuBLAS for a GEMM: plain library
// calls for plain library work); the kernels between them are below.  Results are integers: identical to the host's.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "moira_internal.h"

namespace moira {
namespace {

constexpr uint32_t NONE32 = 0xFFFFFFFFu;

__global__ void __launch_bounds__(256) grp_first_kernel(const uint32_t *__restrict__ label, uint32_t n, uint32_t *first, uint32_t *err)
{
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t l = label[r];
        if (l >= n) { atomicMin(err, r); continue; }
        atomicMin(&first[l], r);
    }
}

__global__ void __launch_bounds__(256) grp_isfirst_kernel(const uint32_t *__restrict__ label, const uint32_t *__restrict__ first, uint32_t n,
                                                          uint32_t *__restrict__ isfirst, uint32_t *__restrict__ iota)
{
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t l = label[r];
        isfirst[r] = (l < n && first[l] == r) ? 1u : 0u;
        iota[r] = r;
    }
}

__global__ void __launch_bounds__(256) grp_gid_kernel(const uint32_t *__restrict__ label, const uint32_t *__restrict__ first,
                                                      const uint32_t *__restrict__ gidx, uint32_t n, uint32_t *__restrict__ g, uint32_t *size)
{
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t l = label[r];
        const uint32_t gg = l < n ? gidx[first[l]] : 0u;
        g[r] = gg;
        atomicAdd(&size[gg], 1u);
    }
}

__global__ void __launch_bounds__(256) grp_gather_kernel(const double *__restrict__ ee, const uint32_t *__restrict__ r_sorted, uint32_t n,
                                                         double *__restrict__ ee_sorted)
{
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) ee_sorted[k] = ee[r_sorted[k]];
}

// brk[k] = 1 where the read is the first of its group or its ee is strictly below every earlier one's of the group
__global__ void __launch_bounds__(256) grp_breaker_kernel(const uint32_t *__restrict__ g_sorted, const uint32_t *__restrict__ mstart,
                                                          const double *__restrict__ ee_sorted, const double *__restrict__ minpref, uint32_t n,
                                                          uint32_t *__restrict__ brk)
{
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t s = mstart[g_sorted[k]];
        brk[k] = (k == s || ee_sorted[k] < minpref[k - 1]) ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(256) grp_place_kernel(const uint32_t *__restrict__ g_sorted, const uint32_t *__restrict__ r_sorted,
                                                        const uint32_t *__restrict__ mstart, const uint32_t *__restrict__ brk,
                                                        const uint32_t *__restrict__ brk_incl, uint32_t n, uint32_t *__restrict__ members,
                                                        uint32_t *__restrict__ rep)
{
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t g = g_sorted[k];
        const uint32_t s = mstart[g], e = mstart[g + 1];
        const uint32_t bg = brk_incl[e - 1], bi = brk_incl[k];
        const uint32_t pos = brk[k] ? s + (bg - bi) : s + bg + (k - s) - bi;
        members[pos] = r_sorted[k];
        if (brk[k] && bi == bg) rep[g] = r_sorted[k];
    }
}

__global__ void __launch_bounds__(256) grp_sizekey_kernel(const uint32_t *__restrict__ size, uint32_t G, uint32_t *__restrict__ key,
                                                          uint32_t *__restrict__ iota)
{
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
        key[g] = ~size[g];      // ascending ~size == descending size; the radix sort is stable: ties keep first-appearance order
        iota[g] = g;
    }
}

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

size_t groups_device_bytes(uint64_t n)
{
    // 17 uint32 arrays of n + 1, 3 double arrays of n, CUB temporary storage (queried generously: 2 x 4n + 16 MB)
    return 17 * al((n + 1) * 4) + 3 * al(n * 8) + al(8 * n + (16u << 20));
}
size_t groups_host_bytes(uint64_t n) { return 6 * al((n + 1) * 4); }

// d_buf: groups_device_bytes(n) bytes of device memory; h_pinned: groups_host_bytes(n) bytes of pinned host memory.
// labels / ee: host arrays of n (device arrays when on_device).  The uint32 results land in h_pinned: [g_of_read n][rep G][size G][mstart G+1][members n][order G]
// at the offsets returned in out_off[6].
int groups_from_labels_device(int sm_count, cudaStream_t stream, uint8_t *d_buf, uint8_t *h_pinned, const uint32_t *labels, const double *ee,
                              int on_device, uint64_t n64, uint64_t *n_groups_out, size_t out_off[6])
{
    if (n64 == 0 || n64 >= 0x7FFFFFF0ull) return fail(MOIRA_ERR_BAD_ARG, "collapse on the device takes 1 .. 2^31 reads");
    const uint32_t n = (uint32_t)n64;
    uint8_t *p = d_buf;
    auto take = [&](size_t bytes) { uint8_t *q = p; p += al(bytes); return q; };
    uint32_t *d_label = (uint32_t *)take(((size_t)n + 1) * 4), *d_first = (uint32_t *)take(((size_t)n + 1) * 4), *d_isfirst = (uint32_t *)take(((size_t)n + 1) * 4);
    uint32_t *d_gidx = (uint32_t *)take(((size_t)n + 1) * 4), *d_g = (uint32_t *)take(((size_t)n + 1) * 4), *d_size = (uint32_t *)take(((size_t)n + 1) * 4);
    uint32_t *d_mstart = (uint32_t *)take(((size_t)n + 1) * 4), *d_iota = (uint32_t *)take(((size_t)n + 1) * 4), *d_gs = (uint32_t *)take(((size_t)n + 1) * 4);
    uint32_t *d_rs = (uint32_t *)take(((size_t)n + 1) * 4), *d_brk = (uint32_t *)take(((size_t)n + 1) * 4), *d_bi = (uint32_t *)take(((size_t)n + 1) * 4);
    uint32_t *d_members = (uint32_t *)take(((size_t)n + 1) * 4), *d_rep = (uint32_t *)take(((size_t)n + 1) * 4), *d_key = (uint32_t *)take(((size_t)n + 1) * 4);
    uint32_t *d_ks = (uint32_t *)take(((size_t)n + 1) * 4), *d_order = (uint32_t *)take(((size_t)n + 1) * 4);
    double *d_ee = (double *)take((size_t)n * 8), *d_ees = (double *)take((size_t)n * 8), *d_minp = (double *)take((size_t)n * 8);
    void *d_tmp = p;
    size_t tmp_bytes = al(8 * (size_t)n + (16u << 20));

    const int grid = sm_count * 8;
#define CK(x) do { if ((x) != cudaSuccess) return fail(MOIRA_ERR_CUDA, "collapse on the device: %s", cudaGetErrorString(cudaGetLastError())); } while (0)
    if (on_device) {   // labels / ee are device arrays already (the filter's and the dereplication's own outputs)
        d_label = const_cast<uint32_t *>(labels);
        d_ee = const_cast<double *>(ee);
    } else {
        CK(cudaMemcpyAsync(d_label, labels, (size_t)n * 4, cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_ee, ee, (size_t)n * 8, cudaMemcpyHostToDevice, stream));
    }
    CK(cudaMemsetAsync(d_first, 0xFF, (size_t)(n + 1) * 4, stream));
    CK(cudaMemsetAsync(d_size, 0, (size_t)(n + 1) * 4, stream));
    uint32_t *h_err = (uint32_t *)h_pinned;
    CK(cudaMemsetAsync(d_key, 0xFF, 4, stream));   // error word: smallest read index with a label that is not a read index
    grp_first_kernel<<<grid, 256, 0, stream>>>(d_label, n, d_first, d_key);
    grp_isfirst_kernel<<<grid, 256, 0, stream>>>(d_label, d_first, n, d_isfirst, d_iota);
    size_t need = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, need, d_isfirst, d_gidx, (int)n + 1, stream));
    if (need > tmp_bytes) return fail(MOIRA_ERR_NOMEM, "scan scratch");
    CK(cudaMemsetAsync(d_isfirst + n, 0, 4, stream));
    need = tmp_bytes;
    CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, d_isfirst, d_gidx, (int)n + 1, stream));   // gidx[n] = G
    CK(cudaMemcpyAsync(h_err, d_key, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(h_err + 1, d_gidx + n, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (h_err[0] != NONE32) return fail(MOIRA_ERR_BAD_ARG, "label of read %u is not a read index", h_err[0]);
    const uint32_t G = h_err[1];
    *n_groups_out = G;
    grp_gid_kernel<<<grid, 256, 0, stream>>>(d_label, d_first, d_gidx, n, d_g, d_size);
    need = tmp_bytes;
    CK(cub::DeviceScan::ExclusiveSum(d_tmp, need, d_size, d_mstart, (int)G + 1, stream));
    int bits = 1;
    while (bits < 32 && (1ull << bits) < (uint64_t)G + 1) bits++;
    need = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, d_g, d_gs, d_iota, d_rs, (int)n, 0, bits, stream));
    if (need > tmp_bytes) return fail(MOIRA_ERR_NOMEM, "sort scratch");
    CK(cub::DeviceRadixSort::SortPairs(d_tmp, need, d_g, d_gs, d_iota, d_rs, (int)n, 0, bits, stream));
    grp_gather_kernel<<<grid, 256, 0, stream>>>(d_ee, d_rs, n, d_ees);
    need = 0;
    CK(cub::DeviceScan::InclusiveScanByKey(nullptr, need, d_gs, d_ees, d_minp, cub::Min(), (int)n, cub::Equality(), stream));
    if (need > tmp_bytes) return fail(MOIRA_ERR_NOMEM, "scan-by-key scratch");
    CK(cub::DeviceScan::InclusiveScanByKey(d_tmp, need, d_gs, d_ees, d_minp, cub::Min(), (int)n, cub::Equality(), stream));
    grp_breaker_kernel<<<grid, 256, 0, stream>>>(d_gs, d_mstart, d_ees, d_minp, n, d_brk);
    need = 0;
    CK(cub::DeviceScan::InclusiveSumByKey(nullptr, need, d_gs, d_brk, d_bi, (int)n, cub::Equality(), stream));
    if (need > tmp_bytes) return fail(MOIRA_ERR_NOMEM, "scan-by-key scratch");
    CK(cub::DeviceScan::InclusiveSumByKey(d_tmp, need, d_gs, d_brk, d_bi, (int)n, cub::Equality(), stream));
    grp_place_kernel<<<grid, 256, 0, stream>>>(d_gs, d_rs, d_mstart, d_brk, d_bi, n, d_members, d_rep);
    grp_sizekey_kernel<<<grid, 256, 0, stream>>>(d_size, G, d_key, d_iota);
    need = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, d_key, d_ks, d_iota, d_order, (int)G, 0, 32, stream));
    if (need > tmp_bytes) return fail(MOIRA_ERR_NOMEM, "sort scratch");
    if (G) CK(cub::DeviceRadixSort::SortPairs(d_tmp, need, d_key, d_ks, d_iota, d_order, (int)G, 0, 32, stream));
    if (cudaGetLastError() != cudaSuccess) return fail(MOIRA_ERR_CUDA, "collapse kernels failed to launch");
    // results -> pinned host memory
    size_t o = 0;
    auto back = [&](int i, const uint32_t *src, size_t count) {
        out_off[i] = o;
        if (count) cudaMemcpyAsync(h_pinned + o, src, count * 4, cudaMemcpyDeviceToHost, stream);
        o += al(count * 4 + 4);
    };
    back(0, d_g, n); back(1, d_rep, G); back(2, d_size, G); back(3, d_mstart, (size_t)G + 1); back(4, d_members, n); back(5, d_order, G);
    CK(cudaStreamSynchronize(stream));
#undef CK
    return MOIRA_OK;
}

}  // namespace moira
