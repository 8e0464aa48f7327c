// Microbenchmark: how fast can one SM accumulate integer co-counts in shared memory?
// Decides the accumulator design of the A^T B kernel (DESIGN.md section "accumulator").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o smem_accum_bench smem_accum_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <math.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__host__ __device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// mode 0: CTA-shared table, atomicAdd on dense counters (random address)
// mode 1: CTA-shared packed hash (key<<12|cnt): LDS probe, atomicAdd if match, atomicCAS if empty
// mode 2: warp-private table, non-atomic LDS/ADD/STS, lanes of one warp-instruction hit distinct slots
// mode 3: CTA-shared table, atomicAdd, lanes of one warp-instruction hit distinct slots
// mode 4: global (L2-resident) dense counters, atomicAdd (RED)
// mode 5: CTA-shared separate key/count arrays: atomicCAS(key) + atomicAdd(cnt)
template <int MODE>
__global__ void accum(const uint32_t* __restrict__ idx, int n_per_cta, int table_slots, uint32_t* __restrict__ gtab,
                      unsigned long long* __restrict__ out) {
    extern __shared__ uint32_t tab[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    if (MODE != 4) for (int i = tid; i < table_slots; i += nt) tab[i] = (MODE == 1 || MODE == 5) ? 0xffffffffu : 0u;
    if (MODE == 5) for (int i = tid; i < table_slots; i += nt) tab[table_slots + i] = 0u;
    __syncthreads();
    const uint32_t* src = idx + (size_t)blockIdx.x * n_per_cta;
    const uint32_t mask = table_slots - 1;
    uint32_t racc = 0;
    if (MODE == 7) {
        uint32_t x = tid * 2654435761u + blockIdx.x;
        for (int i = tid; i < n_per_cta; i += nt) { x = x * 1664525u + 1013904223u; atomicAdd(&tab[(x >> 8) & mask], 1u); }
    } else if (MODE == 8) {
        uint32_t x = tid * 2654435761u + blockIdx.x;
        for (int i = tid; i < n_per_cta; i += nt) { x = x * 1664525u + 1013904223u; uint32_t s = (x >> 8) & mask; s = (s & ~31u) | lane;  atomicAdd(&tab[s], 1u); }
    } else if (MODE == 2) {
        const int wslots = table_slots / nw;
        uint32_t* wt = tab + warp * wslots;
        const uint32_t wmask = wslots - 1;
        uint32_t nxt = src[tid];
        for (int i = tid; i < n_per_cta; i += nt) {
            uint32_t b = nxt;
            if (i + nt < n_per_cta) nxt = src[i + nt];
            // distinct-per-warp slot: rotate base by lane so the 32 lanes never collide
            uint32_t s = ((b & ~31u) + lane) & wmask;
            uint32_t v = wt[s];
            wt[s] = v + 1;
            __syncwarp();
        }
    } else {
        uint32_t nxt[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j) nxt[j] = (tid + j * nt < n_per_cta) ? src[tid + j * nt] : 0;
        for (int i = tid; i < n_per_cta; i += 4 * nt) {
            uint32_t cur[4];
            #pragma unroll
            for (int j = 0; j < 4; ++j) { cur[j] = nxt[j]; int k = i + (4 + j) * nt; nxt[j] = (k < n_per_cta) ? src[k] : 0; }
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (i + j * nt >= n_per_cta) break;
                uint32_t b = cur[j];
                if (MODE == 0) { atomicAdd(&tab[b & mask], 1u); }
                else if (MODE == 6) { racc += b; }
                else if (MODE == 3) { atomicAdd(&tab[((b & ~31u) + lane) & mask], 1u); }
                else if (MODE == 4) { atomicAdd(&gtab[(size_t)blockIdx.x * table_slots + (b & mask)], 1u); }
                else if (MODE == 1) {
                    uint32_t key = b & 0xfffffu;          // 20-bit key space
                    uint32_t s = mix(key) & mask;
                    while (true) {
                        uint32_t w = tab[s];
                        if ((w >> 12) == key) { atomicAdd(&tab[s], 1u); break; }
                        if (w == 0xffffffffu) {
                            uint32_t old = atomicCAS(&tab[s], 0xffffffffu, (key << 12) | 1u);
                            if (old == 0xffffffffu) break;
                            if ((old >> 12) == key) { atomicAdd(&tab[s], 1u); break; }
                        }
                        s = (s + 1) & mask;
                    }
                } else if (MODE == 5) {
                    uint32_t key = b & 0xfffffu;
                    uint32_t s = mix(key) & mask;
                    while (true) {
                        uint32_t old = atomicCAS(&tab[s], 0xffffffffu, key);
                        if (old == 0xffffffffu || old == key) { atomicAdd(&tab[table_slots + s], 1u); break; }
                        s = (s + 1) & mask;
                    }
                }
            }
        }
    }
    __syncthreads();
    unsigned long long acc = 0;
    if (MODE != 4) for (int i = tid; i < table_slots; i += nt) acc += (MODE == 1) ? (tab[i] == 0xffffffffu ? 0 : (tab[i] & 0xfff)) : (MODE == 5 ? tab[table_slots + i] : tab[i]);
    if (MODE == 4) { acc = 0; for (int i = tid; i < table_slots; i += nt) acc += gtab[(size_t)blockIdx.x * table_slots + i]; }
    if (MODE == 6) { if (tid == 0) tab[0] = 0; __syncthreads(); atomicAdd(&tab[0], racc); __syncthreads(); acc = (tid == 0) ? (unsigned long long)n_per_cta + (tab[0] == 0xdeadbeef) : 0; }
    atomicAdd(out, acc);
}

__global__ void gen(uint32_t* d, size_t n, uint32_t keys) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        d[i] = mix((uint32_t)i * 2654435761u + 12345u) % keys;
}

__global__ void log_bench(const double* __restrict__ x, double* __restrict__ y, int n, int reps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = x[i], acc = 0;
    for (int r = 0; r < reps; ++r) { acc += v * log(v); v += 1.0; }
    y[i] = acc;
}

template <int MODE>
void run(const char* name, int threads, int table_slots, int distinct_keys, int ctas_per_sm) {
    int n_per_cta = (1 << 20) / ctas_per_sm;
    int grid = 148 * ctas_per_sm;
    size_t n = (size_t)grid * n_per_cta;
    uint32_t *d, *gt = nullptr; unsigned long long* out;
    CK(cudaMalloc(&d, n * 4));
    gen<<<4096, 256>>>(d, n, distinct_keys); CK(cudaDeviceSynchronize());
    CK(cudaMalloc(&out, 8)); CK(cudaMemset(out, 0, 8));
    if (MODE == 4) { CK(cudaMalloc(&gt, (size_t)grid * table_slots * 4)); CK(cudaMemset(gt, 0, (size_t)grid * table_slots * 4)); }
    size_t smem = (size_t)table_slots * 4 * (MODE == 5 ? 2 : 1);
    if (MODE == 4) smem = 1024;
    CK(cudaFuncSetAttribute(accum<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    accum<MODE><<<grid, threads, smem>>>(d, n_per_cta, table_slots, gt, out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(out, 0, 8));
    if (MODE == 4) CK(cudaMemset(gt, 0, (size_t)grid * table_slots * 4));
    cudaEventRecord(e0);
    accum<MODE><<<grid, threads, smem>>>(d, n_per_cta, table_slots, gt, out);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long total; CK(cudaMemcpy(&total, out, 8, cudaMemcpyDeviceToHost));
    double rate = (double)n / (ms * 1e-3);
    printf("%-34s thr=%4d slots=%6d keys=%7d cta/sm=%d : %8.3f ms  %8.2f Gprod/s  %6.2f prod/clk/SM@1.9GHz  check=%s\n",
           name, threads, table_slots, distinct_keys, ctas_per_sm, ms, rate * 1e-9, rate / 148 / 1.9e9,
           total == n ? "ok" : "MISMATCH");
    cudaFree(d); cudaFree(out); if (gt) cudaFree(gt);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sms=%d smem/blk optin=%zu clock=%d kHz\n", p.name, p.multiProcessorCount, p.sharedMemPerBlockOptin, p.clockRate);
    // stream read baseline is implicit: 4 B per product
    run<0>("cta-shared dense atomicAdd", 1024, 32768, 32768, 1);
    run<0>("cta-shared dense atomicAdd", 1024, 32768, 1000, 1);
    run<0>("cta-shared dense atomicAdd", 512, 16384, 16384, 2);
    run<0>("cta-shared dense atomicAdd", 256, 8192, 8192, 4);
    run<0>("cta-shared dense atomicAdd", 256, 8192, 8192, 8);
    run<6>("load only (no atomics)", 1024, 32768, 32768, 1);
    run<6>("load only (no atomics)", 256, 8192, 8192, 8);
    run<7>("reg-gen idx atomicAdd random", 1024, 32768, 32768, 1);
    run<7>("reg-gen idx atomicAdd random", 512, 16384, 32768, 4);
    run<7>("reg-gen idx atomicAdd random", 256, 8192, 32768, 8);
    run<8>("reg-gen idx atomicAdd conflict-free", 1024, 32768, 32768, 1);
    run<8>("reg-gen idx atomicAdd conflict-free", 256, 8192, 32768, 8);
    run<3>("cta-shared atomicAdd warp-distinct", 1024, 32768, 32768, 1);
    run<3>("cta-shared atomicAdd warp-distinct", 256, 8192, 8192, 8);
    run<1>("cta-shared packed hash", 1024, 32768, 12000, 1);
    run<1>("cta-shared packed hash", 1024, 32768, 1000, 1);
    run<1>("cta-shared packed hash", 256, 8192, 3000, 8);
    run<5>("cta-shared key+cnt CAS hash", 1024, 16384, 6000, 1);
    run<5>("cta-shared key+cnt CAS hash", 256, 4096, 1500, 8);
    run<2>("warp-private non-atomic", 1024, 32768, 32768, 1);
    run<2>("warp-private non-atomic", 1024, 32768, 32768, 2);
    run<2>("warp-private non-atomic", 512, 16384, 16384, 4);
    run<4>("global(L2) dense atomicAdd", 1024, 65536, 65536, 1);
    run<4>("global(L2) dense atomicAdd", 1024, 65536, 2000, 1);
    run<4>("global(L2) dense atomicAdd", 512, 65536, 65536, 4);
    // fp64 log throughput
    {
        int n = 148 * 2048 * 8, reps = 64;
        double *x, *y; CK(cudaMalloc(&x, n * 8)); CK(cudaMalloc(&y, n * 8));
        std::vector<double> hx(n); for (int i = 0; i < n; ++i) hx[i] = 1.0 + (i % 100000);
        CK(cudaMemcpy(x, hx.data(), n * 8, cudaMemcpyHostToDevice));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        log_bench<<<n / 256, 256>>>(x, y, n, reps); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); log_bench<<<n / 256, 256>>>(x, y, n, reps); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double rate = (double)n * reps / (ms * 1e-3);
        printf("fp64 x*log(x): %.3f ms, %.2f G xlogx/s, %.3f per clk per SM\n", ms, rate * 1e-9, rate / 148 / 1.9e9);
    }
    return 0;
}
