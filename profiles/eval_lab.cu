// eval_lab.cu -- stand-alone bench of batch-evaluator variants (profiles/ tool, not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I pokerl_b200/csrc -o profiles/_bin/eval_lab profiles/eval_lab.cu
//   profiles/_bin/eval_lab [log2 hands] [cards per hand, 0 = mixed]   (profiles/_bin/ is git-ignored and travels with gpurun)
// Every variant's output is compared with eval_mask52 (the bitmask evaluator) on all hands.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "pkb_eval.cuh"

using namespace pkb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct FastTables {
    const uint32_t* dense;  // [65536]
    const uint32_t* klo;    // [128]
    const uint32_t* khi;    // [64]
    const uint8_t* disp;    // [16384]
    const uint16_t* dense16;  // [65536] slot -> class
    const uint32_t* vals;     // [nvals] class -> packed value
    uint32_t nvals;
};

__device__ __forceinline__ uint64_t splitmix(uint64_t& s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void random_hands(uint64_t n, uint64_t* masks, int cards)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = i * 0x2545F4914F6CDD1Dull + 12345, m = 0;
        int k = cards;
        if (cards == 0) k = (int)(splitmix(s) % 8);  // mixed sizes
        while (popc64(m) < k) m |= 1ull << (splitmix(s) % 52);
        masks[i] = m;
    }
}
__global__ void reference_kernel(const uint64_t* masks, uint64_t n, uint32_t* out)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = eval_mask52(masks[i]);
}
__global__ void diff_kernel(const uint32_t* a, const uint32_t* b, uint64_t n, unsigned long long* bad)
{
    unsigned long long c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) c += a[i] != b[i];
    if (c) atomicAdd(bad, c);
}

__device__ __noinline__ uint32_t eval_sel_call(uint64_t m) { return eval_mask52_sel(m); }

__device__ __forceinline__ uint32_t lds32(const void* base, uint32_t off) { return *reinterpret_cast<const uint32_t*>(static_cast<const char*>(base) + off); }

// the common case on the device: byte offsets into the per-lane copies (entry x of lane l at x * 128 + l * 4)
template <bool SD>
__device__ __forceinline__ uint32_t fast7(uint32_t lo, uint32_t hi, const uint32_t* s_klo, const uint32_t* s_khi, uint32_t lane4,
                                          const uint8_t* s_disp, const uint32_t* __restrict__ dense, const uint16_t* s_dense16, const uint32_t* s_vals, uint32_t& rare)
{
    const uint32_t mid = __funnelshift_r(lo, hi, 19);
    const uint32_t s0 = lds32(s_klo, ((lo << 7) & 0x3F80u) | lane4) + lds32(s_khi, (lo & 0x1F80u) | lane4);
    const uint32_t s1 = lds32(s_klo, ((lo >> 6) & 0x3F80u) | lane4) + lds32(s_khi, ((lo >> 13) & 0x1F80u) | lane4);
    const uint32_t s2 = lds32(s_klo, (mid & 0x3F80u) | lane4) + lds32(s_khi, ((hi << 6) & 0x1F80u) | lane4);
    const uint32_t s3 = lds32(s_klo, (hi & 0x3F80u) | lane4) + lds32(s_khi, ((hi >> 7) & 0x1F80u) | lane4);
    const uint32_t sum = s0 + s1 + s2 + s3;
    rare = ((s0 | s1 | s2 | s3) & EVAL_FAST_FLUSH_BIT) | ((sum ^ EVAL_FAST_SEVEN) >> 24);
    const uint32_t d = s_disp[fast_bucket(sum)];
    if (SD) return s_vals[s_dense16[fast_slot(sum, d)]];
    return __ldg(dense + fast_slot(sum, d));
}

// MODE 0: rare hands queued per warp and ranked 32 at a time; MODE 1: ranked in place (divergent)
template <int THREADS, int MINB, int HPT, int MODE, bool SD>
__global__ void __launch_bounds__(THREADS, MINB) eval_fast_kernel(const uint64_t* __restrict__ masks, uint64_t n, uint32_t* __restrict__ out, FastTables t)
{
    constexpr int WARPS = THREADS / 32, QCAP = 32 * HPT + 32;  // power of two for HPT = 1 (64), 3 (128) .. use mask below
    constexpr int QSIZE = QCAP <= 64 ? 64 : QCAP <= 128 ? 128 : 256;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* s_klo = reinterpret_cast<uint32_t*>(smem);                       // [128][32]
    uint32_t* s_khi = s_klo + 128 * 32;                                        // [64][32]
    uint8_t* s_disp = reinterpret_cast<uint8_t*>(s_khi + 64 * 32);             // [16384]
    uint32_t(*s_queue)[QSIZE] = reinterpret_cast<uint32_t(*)[QSIZE]>(s_disp + EVAL_FAST_BUCKETS);
    uint16_t* s_dense16 = reinterpret_cast<uint16_t*>(s_queue + (MODE == 0 ? WARPS : 1));  // [65536] (SD)
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_dense16 + EVAL_DENSE_SIZE);              // [nvals]  (SD)
    for (int i = threadIdx.x; i < 128 * 32; i += THREADS) s_klo[i] = t.klo[i >> 5];
    for (int i = threadIdx.x; i < 64 * 32; i += THREADS) s_khi[i] = t.khi[i >> 5];
    for (int i = threadIdx.x; i < (int)EVAL_FAST_BUCKETS / 16; i += THREADS) reinterpret_cast<uint4*>(s_disp)[i] = reinterpret_cast<const uint4*>(t.disp)[i];
    if (SD) {
        for (int i = threadIdx.x; i < (int)EVAL_DENSE_SIZE / 8; i += THREADS) reinterpret_cast<uint4*>(s_dense16)[i] = reinterpret_cast<const uint4*>(t.dense16)[i];
        for (int i = threadIdx.x; i < (int)t.nvals; i += THREADS) s_vals[i] = t.vals[i];
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5, lane4 = lane * 4, lt = (1u << lane) - 1u;
    const uint64_t trips = n / (32 * HPT), wstride = (uint64_t)gridDim.x * WARPS;
    uint32_t head = 0, tail = 0;
    uint32_t* q = s_queue[MODE == 0 ? w : 0];
    for (uint64_t trip = (uint64_t)blockIdx.x * WARPS + w; trip < trips; trip += wstride) {
        const uint64_t pair0 = trip * (16 * HPT) + lane;  // pairs of this trip: pair0 + 32 j
        ulonglong2 m[HPT / 2];
#pragma unroll
        for (int j = 0; j < HPT / 2; ++j) m[j] = reinterpret_cast<const ulonglong2*>(masks)[pair0 + 32 * j];
        uint32_t r[HPT], rare[HPT];
#pragma unroll
        for (int j = 0; j < HPT / 2; ++j) {
            r[2 * j] = fast7<SD>((uint32_t)m[j].x, (uint32_t)(m[j].x >> 32), s_klo, s_khi, lane4, s_disp, t.dense, s_dense16, s_vals, rare[2 * j]);
            r[2 * j + 1] = fast7<SD>((uint32_t)m[j].y, (uint32_t)(m[j].y >> 32), s_klo, s_khi, lane4, s_disp, t.dense, s_dense16, s_vals, rare[2 * j + 1]);
        }
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < HPT; ++j)
                if (rare[j]) r[j] = eval_sel_call(j & 1 ? m[j / 2].y : m[j / 2].x);
        }
#pragma unroll
        for (int j = 0; j < HPT / 2; ++j) reinterpret_cast<uint2*>(out)[pair0 + 32 * j] = make_uint2(r[2 * j], r[2 * j + 1]);
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < HPT; ++j) {
                const uint32_t bal = __ballot_sync(0xffffffffu, rare[j] != 0);
                if (rare[j]) q[(tail + __popc(bal & lt)) & (QSIZE - 1)] = (uint32_t)(2 * (pair0 + 32 * (j / 2)) + (j & 1));
                tail += __popc(bal);
            }
            while (tail - head >= 32) {
                __syncwarp();
                const uint32_t idx = q[(head + lane) & (QSIZE - 1)];
                head += 32;
                out[idx] = eval_mask52_sel(masks[idx]);
            }
        }
    }
    if (MODE == 0) {
        __syncwarp();
        if (lane < tail - head) {
            const uint32_t idx = q[(head + lane) & (QSIZE - 1)];
            out[idx] = eval_mask52_sel(masks[idx]);
        }
    }
}

// the shipped kernel's inner function, for the baseline (tables as in pkb_kernels.cu)
struct HashTables { const uint32_t *dense, *flt, *klo, *khi; const uint16_t* disp; };
__global__ void __launch_bounds__(256, 8) eval_hash3_kernel(const uint64_t* __restrict__ masks, uint64_t n, uint32_t* __restrict__ out, HashTables t)
{
    __shared__ uint32_t s_klo[EVAL_KLO_SIZE], s_khi[EVAL_KHI_SIZE];
    for (int i = threadIdx.x; i < EVAL_KLO_SIZE; i += blockDim.x) s_klo[i] = t.klo[i];
    for (int i = threadIdx.x; i < EVAL_KHI_SIZE; i += blockDim.x) s_khi[i] = t.khi[i];
    __syncthreads();
    const uint64_t pairs = n / 2, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride) {
        const ulonglong2 m = reinterpret_cast<const ulonglong2*>(masks)[i];
        const uint32_t a = popc64(m.x) == 7 ? eval7_hash3(m.x, s_klo, s_khi, t.disp, t.dense, t.flt) : eval_mask52_sel(m.x);
        const uint32_t b = popc64(m.y) == 7 ? eval7_hash3(m.y, s_klo, s_khi, t.disp, t.dense, t.flt) : eval_mask52_sel(m.y);
        reinterpret_cast<uint2*>(out)[i] = make_uint2(a, b);
    }
}

template <typename F>
static void run(const char* name, F launch, const uint32_t* d_ref, uint32_t* d_out, uint64_t n)
{
    unsigned long long* d_bad;
    CK(cudaMalloc(&d_bad, 8));
    CK(cudaMemset(d_bad, 0, 8));
    CK(cudaMemset(d_out, 0xff, n * 4));
    launch();
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    diff_kernel<<<1184, 256>>>(d_ref, d_out, n, d_bad);
    unsigned long long bad = 0;
    CK(cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 3; ++i) launch();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f, tot = 0;
    const int K = 10;
    for (int i = 0; i < K; ++i) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
        tot += ms;
    }
    printf("%-34s mismatches %llu  avg %.1f us  best %.1f us  %.1f G evals/s (%.3f of 6543 GB/s)\n", name, bad, tot / K * 1e3, best * 1e3,
           n / (tot / K) / 1e6, 12.0 * n / (tot / K) / 1e6 / 6543.4);
    cudaFree(d_bad);
}

template <int THREADS, int MINB, int HPT, int MODE, bool SD>
static void run_fast(const uint64_t* d_masks, uint64_t n, uint32_t* d_out, const uint32_t* d_ref, FastTables ft, int sms)
{
    char name[96];
    snprintf(name, sizeof name, "fast T=%d B/SM=%d HPT=%d %s %s", THREADS, MINB, HPT, MODE == 0 ? "queue" : "inline", SD ? "smem-values" : "l1-values");
    constexpr int QCAP = 32 * HPT + 32, QSIZE = QCAP <= 64 ? 64 : QCAP <= 128 ? 128 : 256;
    const int smem = 128 * 32 * 4 + 64 * 32 * 4 + (int)EVAL_FAST_BUCKETS + (MODE == 0 ? THREADS / 32 : 1) * QSIZE * 4 + (SD ? (int)EVAL_DENSE_SIZE * 2 + (int)ft.nvals * 4 : 0);
    CK(cudaFuncSetAttribute(eval_fast_kernel<THREADS, MINB, HPT, MODE, SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    run(name, [&] { eval_fast_kernel<THREADS, MINB, HPT, MODE, SD><<<sms * MINB, THREADS, smem>>>(d_masks, n, d_out, ft); }, d_ref, d_out, n);
}

int main(int argc, char** argv)
{
    const int lg = argc > 1 ? atoi(argv[1]) : 26;
    const int cards = argc > 2 ? atoi(argv[2]) : 7;
    const uint64_t n = 1ull << lg;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    // tables
    std::vector<uint32_t> nf(EVAL_NF_SIZE), dense(EVAL_DENSE_SIZE), dense3(EVAL_DENSE_SIZE), klo(128), khi(64), klo3(128), khi3(64), flt(8192);
    std::vector<uint8_t> disp(EVAL_FAST_BUCKETS);
    std::vector<uint16_t> disp3(EVAL_DISP_SIZE + 1);
    if (eval_nf_table_fill(nf.data(), true) != 49205u) { printf("keys collide\n"); return 1; }
    if (!eval_fast_table_fill(nf.data(), disp.data(), dense.data())) { printf("no fast table\n"); return 1; }
    if (eval_nf_table_fill(nf.data()) != 49205u) { printf("keys collide\n"); return 1; }
    if (!eval_dense_table_fill(nf.data(), disp3.data(), dense3.data())) { printf("no dense table\n"); return 1; }
    for (uint32_t x = 0; x < 128; ++x) { klo[x] = fast_klo_entry(x); klo3[x] = raw_key_sum(x); }
    for (uint32_t x = 0; x < 64; ++x) { khi[x] = fast_khi_entry(x); khi3[x] = raw_key_sum(x << 7); }
    for (uint32_t x = 0; x < 8192; ++x) flt[x] = flush_table_entry(rot_ranks(x));
    auto up = [](const void* p, size_t bytes) { void* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemcpy(d, p, bytes, cudaMemcpyHostToDevice)); return d; };
    std::vector<uint32_t> vals(dense);
    std::sort(vals.begin(), vals.end());
    vals.erase(std::unique(vals.begin(), vals.end()), vals.end());
    std::vector<uint16_t> dense16(EVAL_DENSE_SIZE);
    for (uint32_t i = 0; i < EVAL_DENSE_SIZE; ++i) dense16[i] = (uint16_t)(std::lower_bound(vals.begin(), vals.end(), dense[i]) - vals.begin());
    printf("%zu distinct table values\n", vals.size());
    FastTables ft{(const uint32_t*)up(dense.data(), dense.size() * 4), (const uint32_t*)up(klo.data(), 512), (const uint32_t*)up(khi.data(), 256),
                  (const uint8_t*)up(disp.data(), disp.size()), (const uint16_t*)up(dense16.data(), dense16.size() * 2), (const uint32_t*)up(vals.data(), vals.size() * 4),
                  (uint32_t)vals.size()};
    HashTables ht{(const uint32_t*)up(dense3.data(), dense3.size() * 4), (const uint32_t*)up(flt.data(), flt.size() * 4), (const uint32_t*)up(klo3.data(), 512),
                  (const uint32_t*)up(khi3.data(), 256), (const uint16_t*)up(disp3.data(), disp3.size() * 2)};
    uint64_t* d_masks;
    uint32_t *d_ref, *d_out;
    CK(cudaMalloc(&d_masks, n * 8));
    CK(cudaMalloc(&d_ref, n * 4));
    CK(cudaMalloc(&d_out, n * 4));
    random_hands<<<sms * 8, 256>>>(n, d_masks, cards);
    reference_kernel<<<sms * 8, 256>>>(d_masks, n, d_ref);
    CK(cudaDeviceSynchronize());
    printf("2^%d hands, %d cards (0 = mixed), %d SMs\n", lg, cards, sms);
    run("shipped eval7_hash3 256x8", [&] { eval_hash3_kernel<<<sms * 8, 256>>>(d_masks, n, d_out, ht); }, d_ref, d_out, n);
    run_fast<1024, 1, 2, 0, false>(d_masks, n, d_out, d_ref, ft, sms);
    run_fast<1024, 1, 4, 0, false>(d_masks, n, d_out, d_ref, ft, sms);
    run_fast<1024, 1, 8, 0, false>(d_masks, n, d_out, d_ref, ft, sms);
    run_fast<768, 1, 4, 0, false>(d_masks, n, d_out, d_ref, ft, sms);
    run_fast<512, 1, 8, 0, false>(d_masks, n, d_out, d_ref, ft, sms);
    run_fast<1024, 1, 2, 0, true>(d_masks, n, d_out, d_ref, ft, sms);
    run_fast<1024, 1, 4, 0, true>(d_masks, n, d_out, d_ref, ft, sms);
    run_fast<1024, 1, 8, 0, true>(d_masks, n, d_out, d_ref, ft, sms);
    run_fast<768, 1, 4, 0, true>(d_masks, n, d_out, d_ref, ft, sms);
    run_fast<768, 1, 8, 0, true>(d_masks, n, d_out, d_ref, ft, sms);
    run_fast<512, 1, 4, 0, true>(d_masks, n, d_out, d_ref, ft, sms);
    run_fast<512, 1, 8, 0, true>(d_masks, n, d_out, d_ref, ft, sms);
    return 0;
}
