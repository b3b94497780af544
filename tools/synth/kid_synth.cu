// kid_synth.cu - CPU (threaded) and GPU generators for the synthetic workload of kid_synth.h.
// BENCH / TEST TOOLING, not part of the product library.
#include "kid_synth.h"

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

struct ks_gen {
    ks_config cfg;
    std::vector<uint64_t> prefix;
    std::vector<int32_t> parent;
    int dev = -1;
    uint64_t *d_prefix = nullptr;
    int32_t *d_parent = nullptr;
};

namespace {

__global__ void ks_db_kernel(ks_config c, const uint64_t *prefix, uint64_t i0, uint64_t n,
                             uint64_t *keys, uint32_t *taxa)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        keys[i] = ks_probe_key(c.seed_db, i0 + i);
        taxa[i] = ks_taxon_of(prefix, c.n_taxa, i0 + i);
    }
}

__global__ void ks_reads_kernel(ks_config c, const uint64_t *prefix, const int32_t *parent,
                                uint64_t g0, uint64_t n, uint8_t *seq, uint8_t *qual)
{
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t *s = seq + r * c.stride, *q = qual + r * c.stride;
        ks_make_read(&c, prefix, parent, g0 + r, s, q);
        for (uint32_t i = c.read_len; i < c.stride; i++) { s[i] = 'N'; q[i] = '!'; }
    }
}

template <class F>
void parallel_for(uint64_t n, F f)
{
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (n < 4096) nt = 1;
    std::vector<std::thread> th;
    const uint64_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const uint64_t a = t * per, b = std::min(n, a + per);
        if (a >= b) break;
        th.emplace_back([=] { f(a, b); });
    }
    for (auto &t : th) t.join();
}

int ensure_device(ks_gen *g, int device)
{
    if (g->dev == device) return 0;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    cudaFree(g->d_prefix);
    cudaFree(g->d_parent);
    g->d_prefix = nullptr;
    g->d_parent = nullptr;
    if (cudaMalloc(&g->d_prefix, g->prefix.size() * 8) != cudaSuccess) return -3;
    if (cudaMalloc(&g->d_parent, g->parent.size() * 4) != cudaSuccess) return -3;
    cudaMemcpy(g->d_prefix, g->prefix.data(), g->prefix.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(g->d_parent, g->parent.data(), g->parent.size() * 4, cudaMemcpyHostToDevice);
    g->dev = device;
    return 0;
}

} // namespace

extern "C" {

ks_gen *ks_create(const ks_config *cfg, const uint64_t *prefix, const int32_t *parent)
{
    if (!cfg || !prefix || !parent || cfg->n_taxa < 2 || cfg->stride < cfg->read_len) return nullptr;
    ks_gen *g = new ks_gen;
    g->cfg = *cfg;
    g->prefix.assign(prefix, prefix + cfg->n_taxa + 1);
    g->parent.assign(parent, parent + cfg->n_taxa);
    g->cfg.n_probes = g->prefix.back();
    return g;
}

void ks_free(ks_gen *g)
{
    if (!g) return;
    if (g->dev >= 0) { cudaSetDevice(g->dev); cudaFree(g->d_prefix); cudaFree(g->d_parent); }
    delete g;
}

int ks_db_host(ks_gen *g, uint64_t i0, uint64_t n, uint64_t *keys, uint32_t *taxa)
{
    if (!g || i0 + n > g->cfg.n_probes) return -1;
    const ks_config c = g->cfg;
    const uint64_t *prefix = g->prefix.data();
    parallel_for(n, [=](uint64_t a, uint64_t b) {
        for (uint64_t i = a; i < b; i++) {
            keys[i] = ks_probe_key(c.seed_db, i0 + i);
            taxa[i] = ks_taxon_of(prefix, c.n_taxa, i0 + i);
        }
    });
    return 0;
}

int ks_db_device(ks_gen *g, int device, uint64_t i0, uint64_t n, uint64_t *keys, uint32_t *taxa, void *stream)
{
    if (!g || i0 + n > g->cfg.n_probes) return -1;
    int rc = ensure_device(g, device);
    if (rc) return rc;
    if (n == 0) return 0;
    ks_db_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(g->cfg, g->d_prefix, i0, n, keys, taxa);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int ks_reads_host(ks_gen *g, uint64_t g0, uint64_t n, uint8_t *seq, uint8_t *qual)
{
    if (!g) return -1;
    const ks_config c = g->cfg;
    const uint64_t *prefix = g->prefix.data();
    const int32_t *parent = g->parent.data();
    parallel_for(n, [=](uint64_t a, uint64_t b) {
        for (uint64_t r = a; r < b; r++) {
            uint8_t *s = seq + r * c.stride, *q = qual + r * c.stride;
            ks_make_read(&c, prefix, parent, g0 + r, s, q);
            for (uint32_t i = c.read_len; i < c.stride; i++) { s[i] = 'N'; q[i] = '!'; }
        }
    });
    return 0;
}

int ks_reads_device(ks_gen *g, int device, uint64_t g0, uint64_t n, uint8_t *seq, uint8_t *qual, void *stream)
{
    if (!g) return -1;
    int rc = ensure_device(g, device);
    if (rc) return rc;
    if (n == 0) return 0;
    ks_reads_kernel<<<148 * 8, 128, 0, (cudaStream_t)stream>>>(g->cfg, g->d_prefix, g->d_parent, g0, n, seq, qual);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

} // extern "C"
