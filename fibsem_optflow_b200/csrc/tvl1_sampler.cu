// tvl1_sampler.cu -- subsystem (4): flow -> match sampling on the device.
//
// Replaces the host-side tail of solve_wrapper + random_points in the reference
// (src/optflow.cpp:488-493, 522-572): two full-plane mask downloads, two full-plane flow
// downloads, cv::findNonZero over N pixels and a std::random_shuffle that calls rand() N times
// on the host just to keep the first `npoints` (25) entries.
//
// The result here is IDENTICAL to that code for a given seed, without moving planes or
// shuffling N elements:
//   * libstdc++'s random_shuffle is  for i in 1..N-1: swap(v[i], v[rand() % (i+1)]).
//     Position k < K of the final array holds loc[i*] where i* is the LAST step i > k with
//     rand_i % (i+1) == k (steps below i never touch position i, so it still holds its
//     initial element); if there is none, step i == k moved position j_k there, and the
//     argument repeats from j_k over the steps below k.  So only "max i with j_i == t" for
//     t < K is needed: an embarrassingly parallel search over the rand() stream.
//   * glibc's rand() (TYPE_3) is the linear recurrence r[n] = r[n-31] + r[n-3] mod 2^32,
//     output r[n] >> 1.  Linear means jump-ahead: x^k mod (x^31 - x^28 - 1) over Z/2^32 maps
//     a 31-word window k steps forward, so every thread starts its own chunk of the stream.
//   * loc[] (findNonZero order = row-major) is never materialised: per-row popcounts of the
//     mask + a prefix sum locate the row of the i*-th set pixel, one warp finds the column.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "tvl1_internal.h"

namespace tvl1 {

#define CKS(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(TVL1_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                              \
    } while (0)

static const int RCHUNK = 31 * 64;   // rand() calls per thread
static const int RBITS = 24;         // jump polynomials x^(RCHUNK * 2^b), b < RBITS

struct Poly { uint32_t c[31]; };
struct RandPlan {
    uint32_t base[31];       // s_0..s_30: the 31 state words before the first output
    Poly pw[RBITS];
};

// ---- host: glibc srand() state + jump polynomials

static void glibc_seed_window(long long seed, uint32_t* base)
{
    // glibc srandom_r for TYPE_3 (degree 31, separation 3); seed 0 is replaced by 1;
    // an unseeded process behaves as srand(1)
    unsigned s = seed < 0 ? 1u : (unsigned)seed;
    if (s == 0) s = 1;
    std::vector<uint32_t> r(344);
    int32_t word = (int32_t)s;
    r[0] = (uint32_t)word;
    for (int i = 1; i < 31; i++) {
        const long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        r[i] = (uint32_t)word;
    }
    for (int i = 31; i < 34; i++) r[i] = r[i - 31];
    for (int i = 34; i < 344; i++) r[i] = r[i - 31] + r[i - 3];
    for (int i = 0; i < 31; i++) base[i] = r[313 + i];   // output k is r[344 + k] >> 1
}

static void poly_mul(const Poly& a, const Poly& b, Poly* out)
{
    uint32_t t[61];
    memset(t, 0, sizeof(t));
    for (int i = 0; i < 31; i++)
        for (int j = 0; j < 31; j++) t[i + j] += a.c[i] * b.c[j];
    for (int d = 60; d >= 31; d--) {   // x^31 = x^28 + 1
        t[d - 3] += t[d];
        t[d - 31] += t[d];
        t[d] = 0;
    }
    memcpy(out->c, t, sizeof(out->c));
}

static void poly_xpow(unsigned long long k, Poly* out)
{
    Poly result, sq;
    memset(&result, 0, sizeof(result));
    memset(&sq, 0, sizeof(sq));
    result.c[0] = 1;
    sq.c[1] = 1;
    while (k) {
        if (k & 1) poly_mul(result, sq, &result);
        poly_mul(sq, sq, &sq);
        k >>= 1;
    }
    *out = result;
}

__host__ __device__ static inline void window_apply(const uint32_t* c, uint32_t* win)
{
    // win: s_k..s_{k+30}  ->  s_{k+n}..s_{k+n+30} where c = x^n mod P
    uint32_t ext[61];
    for (int i = 0; i < 31; i++) ext[i] = win[i];
    for (int i = 31; i < 61; i++) ext[i] = ext[i - 31] + ext[i - 3];
    for (int m = 0; m < 31; m++) {
        uint32_t acc = 0;
        for (int j = 0; j < 31; j++) acc += c[j] * ext[m + j];
        win[m] = acc;
    }
}

static void make_plan(long long seed, long long skip, RandPlan* plan)
{
    glibc_seed_window(seed, plan->base);
    if (skip > 0) {   // rand() calls already consumed since srand (the reference's debug mode never reseeds)
        Poly p;
        poly_xpow((unsigned long long)skip, &p);
        window_apply(p.c, plan->base);
    }
    poly_xpow((unsigned long long)RCHUNK, &plan->pw[0]);
    for (int b = 1; b < RBITS; b++) poly_mul(plan->pw[b - 1], plan->pw[b - 1], &plan->pw[b]);
}

// ---- device kernels

// thread t owns shuffle steps i = 1 + t*RCHUNK + n (rand() call number i - 1)
__global__ void __launch_bounds__(128) k_shuffle_search(const RandPlan* __restrict__ plan, long long N, int K,
                                                        int* __restrict__ hit, int* __restrict__ jsmall)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i0 = 1 + t * RCHUNK;
    if (i0 >= N) return;
    uint32_t w[31];
#pragma unroll
    for (int i = 0; i < 31; i++) w[i] = plan->base[i];
    for (int b = 0; b < RBITS; b++)
        if ((t >> b) & 1) window_apply(plan->pw[b].c, w);
    for (int blk = 0; blk < RCHUNK / 31; blk++) {
        const long long ib = i0 + (long long)blk * 31;
        if (ib >= N) break;
#pragma unroll
        for (int m = 0; m < 31; m++) {
            w[m] = w[m] + w[(m + 28) % 31];
            const long long i = ib + m;
            if (i < N) {
                const uint32_t o = w[m] >> 1;
                const uint32_t j = o % (uint32_t)(i + 1);
                if (i < K) jsmall[i] = (int)j;
                else if (j < (uint32_t)K) atomicMax(&hit[j], (int)i);
            }
        }
    }
}

__device__ __forceinline__ unsigned mask4(uint32_t a, uint32_t b)
{
    // per byte: (a > 1) | (b > 1)
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned x = (a >> (8 * k)) & 0xff, y = (b >> (8 * k)) & 0xff;
        m |= (unsigned)((x > 1) | (y > 1)) << k;
    }
    return m;
}

// reference src/optflow.cpp:488-493: mask = threshold(f0,1,1,BINARY) | threshold(f1,1,1,BINARY);
// one warp per row counts its set pixels
__global__ void __launch_bounds__(256) k_mask_rowcount(const uint8_t* __restrict__ f0, size_t p0,
                                                       const uint8_t* __restrict__ f1, size_t p1, int w, int h,
                                                       int* __restrict__ rowcount)
{
    const int row = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= h) return;
    const uint8_t* a = f0 + (size_t)row * p0;
    const uint8_t* b = f1 + (size_t)row * p1;
    int cnt = 0;
    const bool al = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 3) == 0;
    int x = 0;
    if (al) {
        const int w4 = w >> 2;
        for (int q = lane; q < w4; q += 32)
            cnt += __popc(mask4(__ldg(reinterpret_cast<const uint32_t*>(a) + q),
                                __ldg(reinterpret_cast<const uint32_t*>(b) + q)));
        x = w4 << 2;
    }
    for (int xx = x + lane; xx < w; xx += 32) cnt += (a[xx] > 1) | (b[xx] > 1);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, off);
    if (lane == 0) rowcount[row] = cnt;
}

struct PickArgs {
    const uint8_t *f0, *f1;
    size_t p0, p1;
    const float *u, *v;
    size_t pf;   // flow pitch in elements
    int w, h;
    int roi0x, roi0y, roi1x, roi1y;
    float inv_scale;
    int q_is_map;       // the `features` branch of random_points (src/optflow.cpp:544-550): q = map + roi1, no pos term
    int n;
    const int* row;     // row of each target
    const int* rank;    // rank of the target among the row's set pixels
    double* out;        // [5][n]: px, py, qx, qy, then (x, y) packed as two ints per entry
    int* pos;
};

// one warp per sampled point: find the rank-th set pixel of its row, then
// reference src/optflow.cpp:552-556 in fp32
__global__ void __launch_bounds__(128) k_pick_points(const __grid_constant__ PickArgs a)
{
    const int k = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= a.n) return;
    const int y = a.row[k];
    int rk = a.rank[k];
    const uint8_t* r0 = a.f0 + (size_t)y * a.p0;
    const uint8_t* r1 = a.f1 + (size_t)y * a.p1;
    int xfound = -1;
    for (int xb = 0; xb < a.w; xb += 32) {
        const int x = xb + lane;
        const bool set = x < a.w && ((r0[x] > 1) | (r1[x] > 1));
        const unsigned bal = __ballot_sync(0xffffffffu, set);
        const int c = __popc(bal);
        if (rk < c) {
            // the rk-th set bit of bal
            unsigned m = bal;
            for (int q = 0; q < rk; q++) m &= m - 1;
            xfound = xb + __ffs(m) - 1;
            break;
        }
        rk -= c;
    }
    if (lane == 0) {
        const int x = xfound;
        a.pos[2 * k] = x;
        a.pos[2 * k + 1] = y;
        const float fu = a.u[(size_t)y * a.pf + x], fv = a.v[(size_t)y * a.pf + x];
        const float px = (float)(x + a.roi0x) * a.inv_scale;
        const float py = (float)(y + a.roi0y) * a.inv_scale;
        const float qx = a.q_is_map ? (fu + (float)a.roi1x) * a.inv_scale : ((float)(x + a.roi1x) + fu) * a.inv_scale;
        const float qy = a.q_is_map ? (fv + (float)a.roi1y) * a.inv_scale : ((float)(y + a.roi1y) + fv) * a.inv_scale;
        a.out[0 * a.n + k] = (double)px;
        a.out[1 * a.n + k] = (double)py;
        a.out[2 * a.n + k] = (double)qx;
        a.out[3 * a.n + k] = (double)qy;
    }
}

// ---- per-handle scratch

struct SamplerScratch {
    int cap_h = 0, cap_k = 0;
    int* d_rowcount = nullptr;
    int* h_rowcount = nullptr;      // pinned
    RandPlan* d_plan = nullptr;
    int *d_hit = nullptr, *d_jsmall = nullptr, *d_row = nullptr, *d_rank = nullptr, *d_pos = nullptr;
    double* d_out = nullptr;
    int* h_ints = nullptr;          // pinned: hit[K], jsmall[K], pos[2K], row[K], rank[K]
    double* h_out = nullptr;        // pinned: 4K
    RandPlan* h_plan = nullptr;     // pinned
};

void sampler_release(void* p)
{
    SamplerScratch* s = (SamplerScratch*)p;
    if (!s) return;
    cudaFree(s->d_rowcount); cudaFreeHost(s->h_rowcount); cudaFree(s->d_plan);
    cudaFree(s->d_hit); cudaFree(s->d_jsmall); cudaFree(s->d_row); cudaFree(s->d_rank); cudaFree(s->d_pos);
    cudaFree(s->d_out); cudaFreeHost(s->h_ints); cudaFreeHost(s->h_out); cudaFreeHost(s->h_plan);
    delete s;
}

static int scratch_reserve(SamplerScratch* s, int h, int K)
{
    if (h > s->cap_h) {
        cudaFree(s->d_rowcount); cudaFreeHost(s->h_rowcount);
        s->d_rowcount = nullptr; s->h_rowcount = nullptr; s->cap_h = 0;
        CKS(cudaMalloc(&s->d_rowcount, sizeof(int) * (size_t)h));
        CKS(cudaMallocHost(&s->h_rowcount, sizeof(int) * (size_t)h));
        s->cap_h = h;
    }
    if (!s->d_plan) CKS(cudaMalloc(&s->d_plan, sizeof(RandPlan)));
    if (!s->h_plan) CKS(cudaMallocHost(&s->h_plan, sizeof(RandPlan)));
    if (K > s->cap_k) {
        cudaFree(s->d_hit); cudaFree(s->d_jsmall); cudaFree(s->d_row); cudaFree(s->d_rank); cudaFree(s->d_pos);
        cudaFree(s->d_out); cudaFreeHost(s->h_ints); cudaFreeHost(s->h_out);
        s->d_hit = s->d_jsmall = s->d_row = s->d_rank = s->d_pos = nullptr; s->d_out = nullptr;
        s->h_ints = nullptr; s->h_out = nullptr; s->cap_k = 0;
        const size_t k = (size_t)K;
        CKS(cudaMalloc(&s->d_hit, sizeof(int) * k));
        CKS(cudaMalloc(&s->d_jsmall, sizeof(int) * k));
        CKS(cudaMalloc(&s->d_row, sizeof(int) * k));
        CKS(cudaMalloc(&s->d_rank, sizeof(int) * k));
        CKS(cudaMalloc(&s->d_pos, sizeof(int) * 2 * k));
        CKS(cudaMalloc(&s->d_out, sizeof(double) * 4 * k));
        CKS(cudaMallocHost(&s->h_ints, sizeof(int) * 6 * k));
        CKS(cudaMallocHost(&s->h_out, sizeof(double) * 4 * k));
        s->cap_k = K;
    }
    return TVL1_OK;
}

// which initial index ends at each of the first K positions of the shuffled array
static void resolve_origins(long long N, int K, const int* hit, const int* jsmall, std::vector<long long>* origin)
{
    // steps i < K, grouped by target value
    std::vector<std::vector<int>> small((size_t)K);
    for (int i = 1; i < K && i < N; i++)
        if (jsmall[i] < K) small[(size_t)jsmall[i]].push_back(i);   // ascending i
    origin->assign((size_t)K, 0);
    for (int k = 0; k < K; k++) {
        long long cur = k, lim = N;
        for (;;) {
            // last step i in (cur, lim) with j_i == cur
            long long found = -1;
            if (lim > K && hit[cur] >= K) found = hit[cur];   // steps >= K are only below lim when lim == N
            if (found < 0) {
                const std::vector<int>& v = small[(size_t)cur];
                // largest element < min(lim, K) and > cur
                const long long hi = lim < K ? lim : K;
                auto it = std::lower_bound(v.begin(), v.end(), (int)hi);
                if (it != v.begin()) {
                    const int cand = *(it - 1);
                    if (cand > cur) found = cand;
                }
            }
            if (found >= 0) { (*origin)[(size_t)k] = found; break; }
            if (cur == 0) { (*origin)[(size_t)k] = 0; break; }
            const long long nxt = jsmall[cur];   // step i == cur swaps position cur with j_cur
            if (nxt == cur) { (*origin)[(size_t)k] = cur; break; }
            lim = cur;
            cur = nxt;
        }
    }
}

}  // namespace tvl1

using namespace tvl1;

extern "C" {

int tvl1_glibc_rand(long long seed, long long skip, int n, int* out)
{
    if (!out || n < 0 || skip < 0) return fail(TVL1_ERR_INVALID, "bad argument");
    uint32_t w[31];
    glibc_seed_window(seed, w);
    if (skip > 0) {
        Poly p;
        poly_xpow((unsigned long long)skip, &p);
        window_apply(p.c, w);
    }
    for (int k = 0; k < n; k++) {
        const int m = k % 31;
        w[m] = w[m] + w[(m + 28) % 31];
        out[k] = (int)(w[m] >> 1);
    }
    return TVL1_OK;
}

int tvl1_sample_matches(tvl1_handle* H, const uint8_t* d_frame0, size_t pitch0, const uint8_t* d_frame1,
                        size_t pitch1, const float* d_u, const float* d_v, size_t pitch_flow, int width,
                        int height, int roi0_x, int roi0_y, int roi1_x, int roi1_y, float scale, int npoints,
                        long long seed, double* px, double* py, double* qx, double* qy, double* wgt,
                        int* positions, int* n_out, void* stream)
{
    return tvl1_sample_matches_skip(H, d_frame0, pitch0, d_frame1, pitch1, d_u, d_v, pitch_flow, width, height,
                                    roi0_x, roi0_y, roi1_x, roi1_y, scale, npoints, seed, 0, px, py, qx, qy, wgt,
                                    positions, n_out, nullptr, stream);
}

int tvl1_sample_matches_skip(tvl1_handle* H, const uint8_t* d_frame0, size_t pitch0, const uint8_t* d_frame1,
                             size_t pitch1, const float* d_u, const float* d_v, size_t pitch_flow, int width,
                             int height, int roi0_x, int roi0_y, int roi1_x, int roi1_y, float scale, int npoints,
                             long long seed, long long rand_skip, double* px, double* py, double* qx, double* qy,
                             double* wgt, int* positions, int* n_out, long long* rand_used, void* stream)
{
    return tvl1_sample_matches_ex(H, d_frame0, pitch0, d_frame1, pitch1, d_u, d_v, pitch_flow, width, height, roi0_x, roi0_y,
                                  roi1_x, roi1_y, scale, npoints, seed, rand_skip, 0, px, py, qx, qy, wgt, positions, n_out,
                                  rand_used, stream);
}

int tvl1_sample_matches_ex(tvl1_handle* H, const uint8_t* d_frame0, size_t pitch0, const uint8_t* d_frame1,
                           size_t pitch1, const float* d_u, const float* d_v, size_t pitch_flow, int width,
                           int height, int roi0_x, int roi0_y, int roi1_x, int roi1_y, float scale, int npoints,
                           long long seed, long long rand_skip, int q_is_map, double* px, double* py, double* qx,
                           double* qy, double* wgt, int* positions, int* n_out, long long* rand_used, void* stream)
{
    if (rand_used) *rand_used = 0;
    if (rand_skip < 0) return fail(TVL1_ERR_INVALID, "rand_skip must be >= 0");
    if (!H) return fail(TVL1_ERR_INVALID, "handle is null");
    if (!d_frame0 || !d_frame1 || !d_u || !d_v || !px || !py || !qx || !qy || !wgt || !n_out)
        return fail(TVL1_ERR_INVALID, "null pointer");
    if (width <= 0 || height <= 0 || pitch0 < (size_t)width || pitch1 < (size_t)width ||
        pitch_flow < (size_t)width * 4 || pitch_flow % 4)
        return fail(TVL1_ERR_INVALID, "bad geometry");
    if (!(scale > 0.f)) return fail(TVL1_ERR_INVALID, "scale must be > 0");
    if (npoints < 0 || npoints > (1 << 20)) return fail(TVL1_ERR_INVALID, "npoints out of range");
    CKS(cudaSetDevice(handle_device(H)));
    cudaStream_t st = (cudaStream_t)stream;
    void** slot = handle_sampler_slot(H);
    if (!*slot) *slot = new SamplerScratch();
    SamplerScratch* S = (SamplerScratch*)*slot;
    int rc = scratch_reserve(S, height, npoints > 0 ? npoints : 1);
    if (rc) return rc;
    *n_out = 0;

    // 1. per-row popcount of the mask
    k_mask_rowcount<<<(height + 7) / 8, 256, 0, st>>>(d_frame0, pitch0, d_frame1, pitch1, width, height, S->d_rowcount);
    CKS(cudaGetLastError());
    // (every small transfer below goes through copy_words: pinned host memory read / written by a kernel,
    // so that nothing queues behind a bulk flow download or slice upload on the copy engines)
    if ((rc = copy_words(S->d_rowcount, S->h_rowcount, height, st))) return rc;
    CKS(cudaStreamSynchronize(st));
    std::vector<long long> prefix((size_t)height + 1, 0);
    for (int y = 0; y < height; y++) prefix[(size_t)y + 1] = prefix[(size_t)y] + S->h_rowcount[y];
    const long long N = prefix[(size_t)height];
    if (rand_used) *rand_used = N > 0 ? N - 1 : 0;   // random_shuffle makes N-1 rand() calls
    if (N == 0) {
        // dummy point so that the fields are present (src/optflow.cpp:560-569)
        px[0] = py[0] = qx[0] = qy[0] = -1.0;
        wgt[0] = 0.0;
        *n_out = 1;
        return TVL1_OK;
    }
    const int K = (int)std::min<long long>(npoints, N);
    if (K == 0) return TVL1_OK;
    if (N - 1 > (long long)RCHUNK << RBITS) return fail(TVL1_ERR_UNSUPPORTED, "mask too large for the jump table");

    // 2. search the rand() stream for the steps that decide positions 0..K-1
    make_plan(seed, rand_skip, S->h_plan);   // (the previous call's copy has completed: it synchronised)
    if ((rc = copy_words(S->h_plan, S->d_plan, (int)(sizeof(RandPlan) / 4), st))) return rc;
    CKS(cudaMemsetAsync(S->d_hit, 0xff, sizeof(int) * (size_t)K, st));   // -1
    CKS(cudaMemsetAsync(S->d_jsmall, 0, sizeof(int) * (size_t)K, st));
    if (N > 1) {
        const long long threads = (N - 1 + RCHUNK - 1) / RCHUNK;
        const int blocks = (int)((threads + 127) / 128);
        k_shuffle_search<<<blocks, 128, 0, st>>>(S->d_plan, N, K, S->d_hit, S->d_jsmall);
        CKS(cudaGetLastError());
    }
    int* h_hit = S->h_ints;
    int* h_js = S->h_ints + K;
    if ((rc = copy_words(S->d_hit, h_hit, K, st)) || (rc = copy_words(S->d_jsmall, h_js, K, st))) return rc;
    CKS(cudaStreamSynchronize(st));
    std::vector<long long> origin;
    resolve_origins(N, K, h_hit, h_js, &origin);

    // 3. locate the origin-th set pixel (row by prefix sums, column on the device), sample
    int* rows = S->h_ints + 4 * (size_t)K;
    int* ranks = S->h_ints + 5 * (size_t)K;
    for (int k = 0; k < K; k++) {
        const long long o = origin[(size_t)k];
        const size_t y = (size_t)(std::upper_bound(prefix.begin(), prefix.end(), o) - prefix.begin()) - 1;
        rows[k] = (int)y;
        ranks[k] = (int)(o - prefix[y]);
    }
    if ((rc = copy_words(rows, S->d_row, K, st)) || (rc = copy_words(ranks, S->d_rank, K, st))) return rc;
    PickArgs a;
    a.f0 = d_frame0; a.f1 = d_frame1; a.p0 = pitch0; a.p1 = pitch1;
    a.u = d_u; a.v = d_v; a.pf = pitch_flow / 4; a.w = width; a.h = height;
    a.roi0x = roi0_x; a.roi0y = roi0_y; a.roi1x = roi1_x; a.roi1y = roi1_y;
    a.inv_scale = (float)(1. / scale);   // float inv_scale = 1./scale  (src/optflow.cpp:528)
    a.q_is_map = q_is_map != 0;
    a.n = K; a.row = S->d_row; a.rank = S->d_rank; a.out = S->d_out; a.pos = S->d_pos;
    k_pick_points<<<(K + 3) / 4, 128, 0, st>>>(a);
    CKS(cudaGetLastError());
    int* h_pos = S->h_ints + 2 * (size_t)K;
    if ((rc = copy_words(S->d_out, S->h_out, 8 * K, st)) || (rc = copy_words(S->d_pos, h_pos, 2 * K, st))) return rc;
    CKS(cudaStreamSynchronize(st));
    for (int k = 0; k < K; k++) {
        px[k] = S->h_out[0 * (size_t)K + k];
        py[k] = S->h_out[1 * (size_t)K + k];
        qx[k] = S->h_out[2 * (size_t)K + k];
        qy[k] = S->h_out[3 * (size_t)K + k];
        wgt[k] = 1.0;
        if (positions) { positions[2 * k] = h_pos[2 * k]; positions[2 * k + 1] = h_pos[2 * k + 1]; }
    }
    *n_out = K;
    return TVL1_OK;
}

}  // extern "C"
