// kid_sample.cu - sample-end kernels and the small diagnostic kernels behind the C-ABI.
//
// ucount (newkmer_10nx.cpp:596-603) is "number of distinct canonical k-mers hit in this sample,
// per hit taxon".  Distinct keys are distinct table slots, so it is the per-taxon histogram of
// the per-slot seen bits - no k-mer set is needed.
#include "kid_kernels.cuh"

#include <algorithm>
#include <cstdlib>

#define KID_UCOUNT_THREADS 1024

namespace {

// N_SRC == 0: one local bitmap (src.p[0]); otherwise OR of n_src bitmaps, possibly peer memory
// SMEM: the block counts into a shared-memory histogram (n_taxa ints) and adds its non-zero rows to
// ucount once at the end - a sample's hits pile onto few taxa, and global atomics on one address queue up
template <int LAYOUT, bool MULTI, bool SMEM>
__global__ void __launch_bounds__(KID_UCOUNT_THREADS)
kid_ucount_kernel(const void *__restrict__ slots_, const KidPtrList src, int n_src,
                  uint64_t quad0, uint64_t n_quads, int *ucount, int n_taxa)
{
    extern __shared__ int uhist[];
    if (SMEM) {
        for (int i = threadIdx.x; i < n_taxa; i += blockDim.x) uhist[i] = 0;
        __syncthreads();
    }
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads;
         q += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v;
        if (!MULTI) {
            v = __ldg(reinterpret_cast<const uint4 *>(src.p[0]) + quad0 + q);
        } else {
            v = make_uint4(0, 0, 0, 0);
            for (int k = 0; k < n_src; k++) { // peer reads go over NVLink; plain loads, not __ldg
                const uint4 t = reinterpret_cast<const uint4 *>(src.p[k])[quad0 + q];
                v.x |= t.x; v.y |= t.y; v.z |= t.z; v.w |= t.w;
            }
        }
        if ((v.x | v.y | v.z | v.w) == 0) continue;
        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t bits = w[k];
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                const uint64_t slot = ((quad0 + q) * 4 + k) * 32 + b;
                uint32_t taxon;
                if (LAYOUT == KID_LAYOUT_KEYHASH)
                    taxon = (uint32_t)__ldg(static_cast<const uint64_t *>(slots_) + slot) & KID_TAXON_MASK;
                else { // packed sector: taxa live in words 3 and 7
                    const uint64_t sec = slot / KID2_SLOTS_PER_SECTOR;
                    const uint32_t *w = static_cast<const uint32_t *>(slots_) + 8 * sec;
                    taxon = kid2_taxon_of(__ldg(w + 3), __ldg(w + 7), (int)(slot - sec * KID2_SLOTS_PER_SECTOR));
                }
                if (taxon < (uint32_t)n_taxa) atomicAdd((SMEM ? uhist : ucount) + taxon, 1);
            }
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_taxa; i += blockDim.x) {
            const int c = uhist[i];
            if (c) atomicAdd(ucount + i, c);
        }
    }
}

// dst[i] = OR over sources of src[k][word0 + i]; sources may be peer-mapped (NVLink) pointers
__global__ void __launch_bounds__(256)
kid_seen_or_kernel(uint4 *dst, const KidPtrList src, int n_src, uint64_t quad0, uint64_t n_quads)
{
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads;
         q += (uint64_t)gridDim.x * blockDim.x) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int k = 0; k < n_src; k++) {
            const uint4 v = reinterpret_cast<const uint4 *>(src.p[k])[quad0 + q];
            acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w;
        }
        dst[q] = acc;
    }
}

__global__ void kid_lookup_kernel(KidTableView t, const uint64_t *keys, size_t n, uint32_t *out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        uint64_t slot;
        out[i] = kid_lookup_from(t, kid_hash60(keys[i] & KID_MASK60), 0, slot);
    }
}

__global__ void kid_lookup2_kernel(Kid2TableView t, const uint64_t *keys, size_t n, uint32_t *out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[i] & KID_MASK60;
        uint64_t slot;
        out[i] = kid2_lookup_from(t, kid2_home_sector(kid_minimizer_mm(key, t.mm), key, t.line_shift, t.sub_bits), key, 0, slot);
    }
}

__global__ void kid_msca_kernel(KidTreeView t, const int32_t *x, const int32_t *y, size_t n,
                                int32_t *out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = (int32_t)kid_msca(t, (uint32_t)x[i], (uint32_t)y[i]);
}

unsigned grid_for(uint64_t n, unsigned threads)
{
    uint64_t g = (n + threads - 1) / threads;
    const uint64_t cap = 148ull * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}


template <int LAYOUT, bool MULTI>
cudaError_t launch_ucount(const void *slots, const KidPtrList &src, int n_src, uint64_t word0, uint64_t n_words,
                          int *ucount, int n_taxa, cudaStream_t stream)
{
    static const bool want_smem = getenv("KID_UCOUNT_GLOBAL") == nullptr;
    const size_t smem = (size_t)n_taxa * sizeof(int);
    if (want_smem && smem <= KID_SMEM_HIST_MAX_BYTES) {
        auto kern = kid_ucount_kernel<LAYOUT, MULTI, true>;
        if (smem > 48 * 1024) {
            const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        // few, large blocks: every block flushes its histogram once
        const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(148ull * 2, (n_words / 4 + KID_UCOUNT_THREADS - 1) / KID_UCOUNT_THREADS));
        kern<<<blocks, KID_UCOUNT_THREADS, smem, stream>>>(slots, src, n_src, word0 / 4, n_words / 4, ucount, n_taxa);
    } else {
        kid_ucount_kernel<LAYOUT, MULTI, false><<<grid_for(n_words / 4, KID_UCOUNT_THREADS), KID_UCOUNT_THREADS, 0, stream>>>(
            slots, src, n_src, word0 / 4, n_words / 4, ucount, n_taxa);
    }
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}

} // namespace

cudaError_t kid_launch_ucount(const void *slots, int layout, const uint32_t *seen, uint64_t word0,
                              uint64_t n_words, int *ucount, int n_taxa, cudaStream_t stream)
{
    if (n_words == 0) return cudaSuccess;
    KidPtrList l;
    for (int i = 0; i < KID_MAX_OR_SOURCES; i++) l.p[i] = i == 0 ? seen : nullptr;
    return layout == KID_LAYOUT_KEYHASH ? launch_ucount<KID_LAYOUT_KEYHASH, false>(slots, l, 1, word0, n_words, ucount, n_taxa, stream)
                                        : launch_ucount<KID_LAYOUT_MINIMIZER, false>(slots, l, 1, word0, n_words, ucount, n_taxa, stream);
}

cudaError_t kid_launch_ucount_or(const void *slots, int layout, const KidPtrList &src, int n_src, uint64_t word0,
                                 uint64_t n_words, int *ucount, int n_taxa, cudaStream_t stream)
{
    if (n_words == 0) return cudaSuccess;
    return layout == KID_LAYOUT_KEYHASH ? launch_ucount<KID_LAYOUT_KEYHASH, true>(slots, src, n_src, word0, n_words, ucount, n_taxa, stream)
                                        : launch_ucount<KID_LAYOUT_MINIMIZER, true>(slots, src, n_src, word0, n_words, ucount, n_taxa, stream);
}

cudaError_t kid_launch_seen_or(uint32_t *dst, const KidPtrList &src, int n_src, uint64_t word0,
                               uint64_t n_words, cudaStream_t stream)
{
    if (n_words == 0) return cudaSuccess;
    kid_seen_or_kernel<<<grid_for(n_words / 4, 256), 256, 0, stream>>>(
        reinterpret_cast<uint4 *>(dst), src, n_src, word0 / 4, n_words / 4);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t kid_launch_lookup(const KidTableView &t, const uint64_t *keys, size_t n, uint32_t *out,
                              cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    kid_lookup_kernel<<<grid_for(n, 256), 256, 0, stream>>>(t, keys, n, out);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t kid_launch_lookup2(const Kid2TableView &t, const uint64_t *keys, size_t n, uint32_t *out,
                               cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    kid_lookup2_kernel<<<grid_for(n, 256), 256, 0, stream>>>(t, keys, n, out);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}

cudaError_t kid_launch_msca(const KidTreeView &t, const int32_t *x, const int32_t *y, size_t n,
                            int32_t *out, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    kid_msca_kernel<<<grid_for(n, 256), 256, 0, stream>>>(t, x, y, n, out);
    KID_COUNT_LAUNCH();
    return cudaGetLastError();
}
