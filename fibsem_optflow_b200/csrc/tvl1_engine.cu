// tvl1_engine.cu -- per-device solver handle, the pair-level driver loop and the C ABI
// declared in include/tvl1_b200.h.  Replaces TVL1_solve (reference src/optflow.cpp:516-520),
// which builds a fresh cv::cuda::OpticalFlowDual_TVL1 -- and all its buffers -- per call: here
// the handle keeps one arena per device and re-uses it across pairs of the same size.
//
// There is no CPU fallback: every entry point runs CUDA kernels or fails with a status.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "tvl1_kernels.cuh"
#include "tvl1_internal.h"

namespace tvl1 {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(TVL1_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                              \
    } while (0)

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- cubic table (A.4)

static void cubic_coeffs(float x, float* c)
{
    // Keys bicubic, A = -0.75, evaluated in fp32 exactly as OpenCV's interpolateCubic does
    const float A = -0.75f;
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
}

static int upload_cubic_table(int device)
{
    static bool done[64] = {false};
    if (device >= 0 && device < 64 && done[device]) return TVL1_OK;
    float tab[128];
    const float scale = 1.f / 32;
    for (int i = 0; i < 32; i++) cubic_coeffs(i * scale, tab + 4 * i);
    CK(cudaMemcpyToSymbol(c_cubic_tab, tab, sizeof(tab)));
    if (device >= 0 && device < 64) done[device] = true;
    return TVL1_OK;
}

int copy_words(const void* src, void* dst, int n32, cudaStream_t st)
{
    if (n32 <= 0) return TVL1_OK;
    const int blocks = n32 <= 256 ? 1 : (n32 + 1023) / 1024 < 32 ? (n32 + 1023) / 1024 : 32;
    k_copy_words<<<blocks, n32 < 256 ? 32 : 256, 0, st>>>((const int*)src, (int*)dst, n32);
    CK(cudaGetLastError());
    return TVL1_OK;
}

// ---------------------------------------------------------------- tensor maps (TMA)

// cuTensorMapEncodeTiled through the runtime's driver entry point: the library keeps linking the CUDA
// runtime only (statically), never libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// an fp32 plane {pitch x h} as a 2-D tensor whose box is one staged tile (box_w x box_h elements, dense rows)
int make_plane_map(CUtensorMap* tm, const float* base, int pitch, int h, int box_w, int box_h)
{
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return fail(TVL1_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)h};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}, es[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TVL1_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %dx%d plane", (int)r, pitch, h);
    return TVL1_OK;
}

// n fp32 planes {pitch x h} at equal distances as a 3-D tensor {x, y, plane}; the box is one 128-px row of every plane
static int make_section_map(CUtensorMap* tm, const float* base, int pitch, int h, int n, size_t plane_stride_bytes)
{
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return fail(TVL1_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(float), (cuuint64_t)plane_stride_bytes};
    const cuuint32_t box[3] = {128u, 1u, (cuuint32_t)n}, es[3] = {1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(TVL1_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for %d planes of %dx%d", (int)r, n, pitch, h);
    return TVL1_OK;
}

// the tensor maps the two-iteration pass fills its ring through (IterMaps), from the plane pointers of `a`:
// the planes of each section must sit at equal, 16-byte aligned distances
static int make_section(CUtensorMap* tm, const float* const* planes, int n, int pitch, int h)
{
    const ptrdiff_t d = (const char*)planes[1] - (const char*)planes[0];
    bool even = d > 0 && d % 16 == 0 && (size_t)d >= (size_t)pitch * h * sizeof(float) && !((uintptr_t)planes[0] & 15);
    for (int k = 2; k < n && even; k++) even = (const char*)planes[k] - (const char*)planes[k - 1] == d;
    if (!even) return fail(TVL1_ERR_INVALID, "the planes of a ring section must sit at equal 16-byte aligned distances");
    return make_section_map(tm, planes[0], pitch, h, n, (size_t)d);
}

int make_iter_maps(IterArgs& a)
{
    int rc;
    const float* c[3] = {a.I1wx, a.I1wy, a.rho_c};
    if ((rc = make_section(&a.tm.c, c, 3, a.pitch, a.h))) return rc;
    for (int t = 0; t < 2; t++) {
        const float* u[2] = {a.u1[t], a.u2[t]};
        const float* p[4] = {a.p11[t], a.p12[t], a.p21[t], a.p22[t]};
        if ((rc = make_section(&a.tm.u[t], u, 2, a.pitch, a.h))) return rc;
        if ((rc = make_section(&a.tm.p[t], p, 4, a.pitch, a.h))) return rc;
    }
    return TVL1_OK;
}

// ---------------------------------------------------------------- pyramid geometry (A.2)

static int scaled_size(int n, double f) { return (int)lrint((double)n * f); }   // round half even

int pyramid_sizes(int w, int h, int nscales, double scale_step, int* ws, int* hs)
{
    if (nscales > TVL1_MAX_LEVELS) nscales = TVL1_MAX_LEVELS;
    ws[0] = w;
    hs[0] = h;
    int used = nscales;
    for (int s = 1; s < nscales; s++) {
        ws[s] = scaled_size(ws[s - 1], scale_step);
        hs[s] = scaled_size(hs[s - 1], scale_step);
        if (ws[s] < 16 || hs[s] < 16) { used = s; break; }   // that level is built then dropped
    }
    return used;
}

// ---------------------------------------------------------------- launch helpers

static inline dim3 grid2d(int w, int h, dim3 b) { return dim3(cdiv(w, b.x), cdiv(h, b.y)); }

int launch_convert(const uint8_t* src, size_t pitch, int w, int h, float* dst, int dpitch, cudaStream_t st)
{
    dim3 b(32, 8);
    dim3 g(cdiv(cdiv(w, 4), 32), cdiv(h, 8));
    k_convert_u8<<<g, b, 0, st>>>(src, pitch, w, h, dst, dpitch);
    CK(cudaGetLastError());
    return TVL1_OK;
}

// one or two planes (srcB / dstB may be null) resized alike in one launch
int launch_resize2(const float* srcA, const float* srcB, int sw, int sh, int sp, float* dstA, float* dstB, int dw, int dh,
                   int dp, double inv_scale, float mul, int apply_mul, cudaStream_t st)
{
    double inv_x, inv_y;
    if (inv_scale > 0) { inv_x = inv_scale; inv_y = inv_scale; }
    else { inv_x = (double)dw / sw; inv_y = (double)dh / sh; }
    const double scale_x = 1. / inv_x, scale_y = 1. / inv_y;
    dim3 b(32, 8);
    if (inv_scale == 0.5 && !apply_mul) {   // resize(src, Size(), 0.5, 0.5): OpenCV's INTER_AREA fast path
        dim3 gh(cdiv(dw, 32), cdiv(dh, 8), srcB ? 2 : 1);
        k_resize_half<<<gh, b, 0, st>>>(srcA, srcB, sw, sh, sp, dstA, dstB, dw, dh, dp);
        CK(cudaGetLastError());
        return TVL1_OK;
    }
    dim3 g(cdiv(cdiv(dw, 4), 32), cdiv(dh, 8), srcB ? 2 : 1);
    k_resize<<<g, b, 0, st>>>(srcA, srcB, sw, sh, sp, dstA, dstB, dw, dh, dp, scale_x, scale_y, mul, apply_mul);
    CK(cudaGetLastError());
    return TVL1_OK;
}

int launch_resize(const float* src, int sw, int sh, int sp, float* dst, int dw, int dh, int dp,
                  double inv_scale, float mul, int apply_mul, cudaStream_t st)
{
    return launch_resize2(src, nullptr, sw, sh, sp, dst, nullptr, dw, dh, dp, inv_scale, mul, apply_mul, st);
}

int launch_gradient(const float* src, int w, int h, int pitch, float* dx, float* dy, cudaStream_t st)
{
    dim3 b(32, 8);
    k_centered_gradient<<<grid2d(w, h, b), b, 0, st>>>(src, w, h, pitch, dx, dy);
    CK(cudaGetLastError());
    return TVL1_OK;
}

int launch_warp(const WarpArgs& a, cudaStream_t st)
{
    // persistent blocks: one grid of resident blocks walks the tile list
    static int cached[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int& resident = cached[dev & 63];
    if (!resident) {
        int sms = 148, occ = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_warp, TVL1_WP_THREADS, 0) != cudaSuccess || occ < 1) occ = 2;
        resident = sms * occ;
    }
    const long long ntiles = (long long)cdiv(a.w, TVL1_WP_TW) * cdiv(a.h, TVL1_WP_TH);
    dim3 b(32, TVL1_WP_NW);
    k_warp<<<(unsigned)(ntiles < resident ? ntiles : resident), b, 0, st>>>(a);
    CK(cudaGetLastError());
    return TVL1_OK;
}

#ifndef TVL1_ITER_NW
#define TVL1_ITER_NW 4
#endif
static const int ITER_NW = TVL1_ITER_NW;

// resident blocks of the iteration kernels on the current device (SMs x occupancy), queried once
// per device; the fused kernel's shared-memory ring needs the opt-in limit raised first
enum { KI_SMALL = 0, KI_LARGE = 1, KI_FUSED = 2, KI_MULTI = 3, KI_OUTER = 4 };   // k_iterate<.,5>, k_iterate<.,4>, k_iterate2, k_iterate_multi
static const long long ITER_LARGE_PX = 16000000;     // levels at least this large: 4 blocks per SM

static int resident_blocks(int which)
{
    static int cached[5][64] = {};
    int dev = 0, sms = 148, occ = 0;
    cudaGetDevice(&dev);
    int& c = cached[which][dev & 63];
    if (c) return c;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e;
    if (which == KI_FUSED) {
        // the shared-memory ring needs the opt-in limit raised first
        cudaFuncSetAttribute(k_iterate2<ITER_NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, TVL1_RING_BYTES(ITER_NW));
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_iterate2<ITER_NW>, 32 * ITER_NW, TVL1_RING_BYTES(ITER_NW));
    } else if (which == KI_OUTER) {
        cudaFuncSetAttribute(k_outer<ITER_NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, TVL1_RING_BYTES(ITER_NW));
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_outer<ITER_NW>, 32 * ITER_NW, TVL1_RING_BYTES(ITER_NW));
    } else if (which == KI_MULTI) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_iterate_multi<ITER_NW, 4>, 32 * ITER_NW, 0);
    } else if (which == KI_LARGE) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_iterate<ITER_NW, 4>, 32 * ITER_NW, 0);
    } else {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_iterate<ITER_NW, 5>, 32 * ITER_NW, 0);
    }
    if (e != cudaSuccess || occ < 1) occ = (which == KI_FUSED || which == KI_OUTER) ? 3 : 4;
    c = sms * occ;
    return c;
}

// upper bound of the grid of any iteration kernel (blocks resident on the device), for the
// per-block partial sums: SMs x the 32-blocks-per-SM hardware limit, two slabs
size_t iterate_max_blocks()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (size_t)sms * 32 * 2;
}

// Rows per tile and grid size.  A tile is one warp's strip x R rows and every resident warp walks
// the tile list with a grid stride: tall tiles amortise the halo rows (R / (R + halo)), but the
// tile count has to spread evenly over the resident warps (no nearly empty last round).
static int tile_rows_search(int w, int h, int strip, int halo, int rmin, int resident_blocks, int* grid)
{
    const long long ns = cdiv(w, strip);
    const long long slots = (long long)resident_blocks * ITER_NW;
    double best = -1.0;
    int best_r = rmin;
    long long best_g = 1;
    static const int rmax = getenv("TVL1_DEV_RMAX") ? atoi(getenv("TVL1_DEV_RMAX")) : 64;   // developer sweeps only
    for (int R = rmin; R <= rmax; ++R) {
        const long long ntiles = ns * cdiv(h, R);
        const long long G = ntiles < slots ? ntiles : slots;   // warps that get work
        const long long rounds = (ntiles + G - 1) / G;
        double eff = (double)ntiles / (double)(rounds * G) * (double)R / (double)(R + halo);
        if (G < slots) eff *= (double)G / (double)slots;       // not even one tile per warp
        if (eff >= best) { best = eff; best_r = R; best_g = G; }
    }
    static const char* dev_rows[2] = {getenv("TVL1_DEV_ROWS"), getenv("TVL1_DEV_ROWS2")};   // developer sweeps only
    if (const char* e = dev_rows[halo == 1 ? 0 : 1]) {
        best_r = atoi(e);
        const long long ntiles = ns * cdiv(h, best_r);
        best_g = ntiles < slots ? ntiles : slots;
    }
    static const bool verbose = getenv("TVL1_DEV_VERBOSE") != nullptr;
    if (verbose)
        fprintf(stderr, "tile_rows %dx%d strip %d: R=%d warps=%lld tiles=%lld\n", w, h, strip, best_r, best_g, ns * cdiv(h, best_r));
    *grid = (int)((best_g + ITER_NW - 1) / ITER_NW);
    return best_r;
}

// the search runs once per (size, kernel): a solve launches the same few shapes hundreds of times
static int tile_rows(int w, int h, int strip, int halo, int rmin, int resident_blocks, int* grid)
{
    struct Entry { int w, h, strip, resident, rows, grid; };
    static thread_local Entry cache[32];
    static thread_local int used = 0, next = 0;
    for (int i = 0; i < used; i++) {
        const Entry& e = cache[i];
        if (e.w == w && e.h == h && e.strip == strip && e.resident == resident_blocks) { *grid = e.grid; return e.rows; }
    }
    Entry e = {w, h, strip, resident_blocks, 0, 0};
    e.rows = tile_rows_search(w, h, strip, halo, rmin, resident_blocks, &e.grid);
    cache[next] = e;
    next = (next + 1) % 32;
    if (used < 32) used++;
    *grid = e.grid;
    return e.rows;
}

int launch_iterate(IterArgs& a, cudaStream_t st)
{
    int grid = 1;
    const bool large = (long long)a.w * a.h >= ITER_LARGE_PX;
    a.rows = tile_rows(a.w, a.h, TVL1_STRIP, 1, 4, resident_blocks(large ? KI_LARGE : KI_SMALL), &grid);
    dim3 b(32, ITER_NW);
    if (large) k_iterate<ITER_NW, 4><<<grid, b, 0, st>>>(a);
    else k_iterate<ITER_NW, 5><<<grid, b, 0, st>>>(a);
    CK(cudaGetLastError());
    return TVL1_OK;
}

// several iterations in one cooperative launch (small levels); the grid must be co-resident
int launch_iterate_multi(IterArgs& a, cudaStream_t st)
{
    // these levels are bound by latency, not by throughput: the shortest tiles that still give every
    // tile its own warp (one round) make an iteration a chain of R + 1 dependent row steps only
    const int resident = resident_blocks(KI_MULTI);
    const long long slots = (long long)resident * ITER_NW, ns = cdiv(a.w, TVL1_STRIP);
    int grid = 1, R = 0;
    for (int r = 1; r <= 8 && !R; ++r)
        if (ns * cdiv(a.h, r) <= slots) R = r;
    if (R) {
        a.rows = R;
        grid = (int)((ns * cdiv(a.h, R) + ITER_NW - 1) / ITER_NW);
    } else {
        a.rows = tile_rows(a.w, a.h, TVL1_STRIP, 1, 4, resident, &grid);
    }
    dim3 b(32, ITER_NW), g(grid);
    void* args[] = {(void*)&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_iterate_multi<ITER_NW, 4>, g, b, args, 0, st);
    if (e != cudaSuccess) return fail(TVL1_ERR_CUDA, "cooperative launch of k_iterate_multi failed: %s", cudaGetErrorString(e));
    return TVL1_OK;
}

// the inner loop of one outer iteration (fused + single passes) in one cooperative launch
int launch_outer(IterArgs& a, cudaStream_t st)
{
    int grid = 1, g1 = 1;
    const int resident = resident_blocks(KI_OUTER);
    const long long slots = (long long)resident * ITER_NW;
    const long long ns2 = cdiv(a.w, TVL1_STRIP2), ns1 = cdiv(a.w, TVL1_STRIP);
    int R2 = 0, R1 = 0;
    // small levels are bound by latency: the shortest tiles that still give every tile its own warp
    for (int r = 1; r <= 8 && !R2; ++r)
        if (ns2 * cdiv(a.h, r) <= slots) R2 = r;
    for (int r = 1; r <= 8 && !R1; ++r)
        if (ns1 * cdiv(a.h, r) <= slots) R1 = r;
    if (R2) {
        a.rows = R2;
        grid = (int)((ns2 * cdiv(a.h, R2) + ITER_NW - 1) / ITER_NW);
    } else {
        a.rows = tile_rows(a.w, a.h, TVL1_STRIP2, 3, 8, resident, &grid);
    }
    a.rows1 = R1 ? R1 : tile_rows(a.w, a.h, TVL1_STRIP, 1, 4, resident, &g1);   // single passes stride over the same grid
    dim3 b(32, ITER_NW), g(grid);
    void* args[] = {(void*)&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_outer<ITER_NW>, g, b, args, TVL1_RING_BYTES(ITER_NW), st);
    if (e != cudaSuccess) return fail(TVL1_ERR_CUDA, "cooperative launch of k_outer failed: %s", cudaGetErrorString(e));
    return TVL1_OK;
}

int launch_iterate2(IterArgs& a, cudaStream_t st)
{
    int grid = 1;
    a.rows = tile_rows(a.w, a.h, TVL1_STRIP2, 3, 8, resident_blocks(KI_FUSED), &grid);   // 3 halo rows per tile
    dim3 b(32, ITER_NW);
    k_iterate2<ITER_NW><<<grid, b, TVL1_RING_BYTES(ITER_NW), st>>>(a);
    CK(cudaGetLastError());
    return TVL1_OK;
}

// gamma != 0: one inner iteration = estimateU launch + dual-update launch (persistent blocks over row segments)
int launch_gamma_iteration(const GammaArgs& a, cudaStream_t st)
{
    static int cached[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int& resident = cached[dev & 63];
    if (!resident) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        resident = sms * 8;
    }
    const long long ntiles = (long long)cdiv(a.w, 128) * a.h;
    const long long want = (ntiles + ITER_NW - 1) / ITER_NW;
    const unsigned grid = (unsigned)(want < resident ? want : resident);
    dim3 b(32, ITER_NW);
    k_gamma_u<ITER_NW><<<grid, b, 0, st>>>(a);
    k_gamma_p<ITER_NW><<<grid, b, 0, st>>>(a);
    CK(cudaGetLastError());
    return TVL1_OK;
}

int launch_median3(const MedianArgs& a, int planes, cudaStream_t st)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long n = (long long)cdiv(a.w, 4) * a.h * planes;
    const long long want = (n + 255) / 256;
    k_median3<<<(unsigned)(want < sms * 8 ? want : sms * 8), 256, 0, st>>>(a, planes);
    CK(cudaGetLastError());
    return TVL1_OK;
}

int launch_median(const MedianArgs& a, int planes, cudaStream_t st)
{
    // persistent blocks: one grid of resident blocks walks the tile list of both planes
    static int cached[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int& resident = cached[dev & 63];
    if (!resident) {
        int sms = 148, occ = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_median5, 256, 0) != cudaSuccess || occ < 1) occ = 2;
        resident = sms * occ;
    }
    const long long ntiles = (long long)cdiv(a.w, TVL1_MED_TW) * cdiv(a.h, TVL1_MED_TH) * planes;
    dim3 b(32, 8);
    k_median5<<<(unsigned)(ntiles < resident ? ntiles : resident), b, 0, st>>>(a, planes);
    CK(cudaGetLastError());
    return TVL1_OK;
}

// ---------------------------------------------------------------- handle

struct Level {
    alignas(64) CUtensorMap tm_med[2][2];   // median tiles of [u1 | u2][twin] (built with the arena)
    CUtensorMap tm_I1[2];                   // warp windows of I1: RW x 16 and RW x RH boxes
    IterMaps tm_iter;                       // ring rows of the two-iteration pass
    int w, h, pitch;
    float *I0, *I1, *u1, *u2;   // u1/u2: buffer [0] of the twin pair; [1] is shared scratch
};

}  // namespace tvl1

using namespace tvl1;

#define TVL1_STACK_SLOTS 4   // slice slots of tvl1_stack_run: two for the pair being solved, two for the next pair's frames

struct tvl1_handle {
    int device = 0;
    tvl1_params prm;
    int inner = 30, outer = 10;
    bool timing = false;
 bool coop_outer = true;             // fused levels: the inner loop of an outer iteration in ONE cooperative launch
    bool multi_iter = true;             // levels below fused_min_px: all inner iterations of an outer one in ONE cooperative launch
    long long fused_min_px = 0;         // levels at least this large use the two-iteration passes (with the cooperative loop they win at every size)
    // arena
    char* arena = nullptr;
    size_t arena_bytes = 0;
    int cap_w = 0, cap_h = 0, cap_scales = 0;
    double cap_step = 0;
    int nlevels = 0;
    Level lv[TVL1_MAX_LEVELS];
    float *I1wx = nullptr, *I1wy = nullptr, *rho = nullptr;   // warp outputs (I1x, I1y, grad never exist as planes)
    float *u1x = nullptr, *u2x = nullptr;   // twin [1] of u, shared by all levels
    float* p[4][2] = {{nullptr}};           // p11,p12,p21,p22 twins
    // gamma != 0 only: u3 per level, p31, p32 (their own allocation, made on first use)
    char* garena = nullptr;
    float* u3[TVL1_MAX_LEVELS] = {};
    float *p31 = nullptr, *p32 = nullptr;
    Ctrl* d_ctrl = nullptr;
    Ctrl* h_ctrl = nullptr;                 // pinned
    double* d_partials = nullptr;
    size_t partials_cap = 0;
    // staging for the host-buffer entry point
    uint8_t *d_f0 = nullptr, *d_f1 = nullptr;
    float *d_uo = nullptr, *d_vo = nullptr;
    size_t stage_pitch8 = 0;
    int stage_w = 0, stage_h = 0;
    cudaStream_t own_stream = nullptr;
    std::vector<cudaEvent_t> events;
    size_t ev_used = 0;
    // sampler scratch (tvl1_sampler.cu), feature scratch (tvl1_features.cu)
    void* samp = nullptr;
    void* feat = nullptr;
    // stack runner (tvl1_stack_run): 3 slice slots, 2 flow buffers, copy streams
    uint8_t* st_slice[TVL1_STACK_SLOTS] = {};
    uint8_t* st_raw[2] = {nullptr, nullptr};   // raw slices awaiting the device prescale
    size_t st_raw_bytes = 0;
    float* st_flow[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    size_t st_pitch8 = 0;
    int st_w = 0, st_h = 0;
    cudaStream_t st_in = nullptr, st_out = nullptr;
    cudaEvent_t st_up[TVL1_STACK_SLOTS] = {}, st_ready[2] = {nullptr, nullptr}, st_down[2] = {nullptr, nullptr};
};

namespace tvl1 {

int handle_device(const tvl1_handle* h) { return h->device; }
void** handle_sampler_slot(tvl1_handle* h) { return &h->samp; }
void** handle_feature_slot(tvl1_handle* h) { return &h->feat; }

static int get_event(tvl1_handle* H, cudaEvent_t* out)
{
    if (H->ev_used == H->events.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        H->events.push_back(e);
    }
    *out = H->events[H->ev_used++];
    return TVL1_OK;
}

static void release_arena(tvl1_handle* H)
{
    if (H->garena) cudaFree(H->garena);
    H->garena = nullptr;
    if (H->arena) cudaFree(H->arena);
    H->arena = nullptr;
    H->arena_bytes = 0;
    H->cap_w = H->cap_h = 0;
}

static int ensure_capacity(tvl1_handle* H, int w, int h, cudaStream_t st)
{
    if (H->arena && H->cap_w == w && H->cap_h == h && H->cap_scales == H->prm.nscales &&
        H->cap_step == H->prm.scale_step)
        return TVL1_OK;
    release_arena(H);
    int ws[TVL1_MAX_LEVELS + 1], hs[TVL1_MAX_LEVELS + 1];
    const int L = pyramid_sizes(w, h, H->prm.nscales, H->prm.scale_step, ws, hs);
    // carve: 256-byte aligned planes
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    struct Plan { size_t I0, I1, u1, u2; } plan[TVL1_MAX_LEVELS];
    for (int s = 0; s < L; s++) {
        const int pitch = round_up(ws[s], 32);
        const size_t b = (size_t)pitch * hs[s] * sizeof(float);
        plan[s].I0 = carve(b); plan[s].I1 = carve(b); plan[s].u1 = carve(b); plan[s].u2 = carve(b);
    }
    const int pitch0 = round_up(w, 32);
    const size_t b0 = (size_t)pitch0 * h * sizeof(float);
    size_t o_scr[5], o_p[8];
    for (int i = 0; i < 5; i++) o_scr[i] = carve(b0);   // I1wx I1wy rho u1x u2x
    for (int i = 0; i < 8; i++) o_p[i] = carve(b0);
    const size_t nb = iterate_max_blocks();
    const size_t o_part = carve(nb * sizeof(double));
    const size_t o_ctrl = carve(sizeof(Ctrl));
    cudaError_t e = cudaMalloc(&H->arena, off);
    if (e != cudaSuccess) {
        H->arena = nullptr;
        return fail(TVL1_ERR_NOMEM, "cudaMalloc(%zu bytes) for a %dx%d pair failed: %s", off, w, h,
                    cudaGetErrorString(e));
    }
    H->arena_bytes = off;
    for (int s = 0; s < L; s++) {
        H->lv[s].w = ws[s]; H->lv[s].h = hs[s]; H->lv[s].pitch = round_up(ws[s], 32);
        H->lv[s].I0 = (float*)(H->arena + plan[s].I0); H->lv[s].I1 = (float*)(H->arena + plan[s].I1);
        H->lv[s].u1 = (float*)(H->arena + plan[s].u1); H->lv[s].u2 = (float*)(H->arena + plan[s].u2);
    }
    H->nlevels = L;
    float** scr[5] = {&H->I1wx, &H->I1wy, &H->rho, &H->u1x, &H->u2x};
    for (int i = 0; i < 5; i++) *scr[i] = (float*)(H->arena + o_scr[i]);
    for (int i = 0; i < 8; i++) H->p[i / 2][i % 2] = (float*)(H->arena + o_p[i]);
    H->d_partials = (double*)(H->arena + o_part);
    H->partials_cap = nb;
    H->d_ctrl = (Ctrl*)(H->arena + o_ctrl);
    for (int s = 0; s < L; s++) {   // tensor maps of the planes TMA reads, once per arena
        Level& lv = H->lv[s];
        const float* planes[2][2] = {{lv.u1, H->u1x}, {lv.u2, H->u2x}};
        for (int z = 0; z < 2; z++)
            for (int t = 0; t < 2; t++) {
                const int r = make_plane_map(&lv.tm_med[z][t], planes[z][t], lv.pitch, lv.h, TVL1_MED_SW, TVL1_MED_SH);
                if (r) { release_arena(H); return r; }
            }
        for (int k = 0; k < 2; k++) {
            const int r = make_plane_map(&lv.tm_I1[k], lv.I1, lv.pitch, lv.h, TVL1_WP_RW, k == 0 ? 16 : TVL1_WP_RH);
            if (r) { release_arena(H); return r; }
        }
        {
            IterArgs ia;
            ia.I1wx = H->I1wx; ia.I1wy = H->I1wy; ia.rho_c = H->rho;
            ia.u1[0] = lv.u1; ia.u1[1] = H->u1x; ia.u2[0] = lv.u2; ia.u2[1] = H->u2x;
            for (int k = 0; k < 2; k++) { ia.p11[k] = H->p[0][k]; ia.p12[k] = H->p[1][k]; ia.p21[k] = H->p[2][k]; ia.p22[k] = H->p[3][k]; }
            ia.w = lv.w; ia.h = lv.h; ia.pitch = lv.pitch;
            const int r = make_iter_maps(ia);
            if (r) { release_arena(H); return r; }
            lv.tm_iter = ia.tm;
        }
    }
    H->cap_w = w; H->cap_h = h; H->cap_scales = H->prm.nscales; H->cap_step = H->prm.scale_step;
    // pad columns are read (never used) by the vectorised kernels: give them defined contents.  On the
    // solve's own stream: a non-blocking stream does not order itself after the legacy default stream
    CK(cudaMemsetAsync(H->arena, 0, off, st));
    return TVL1_OK;
}

// planes of the third channel (gamma != 0): sized like the arena they accompany, released with it
static int ensure_gamma(tvl1_handle* H, cudaStream_t st)
{
    if (H->prm.gamma == 0.0 || H->garena) return TVL1_OK;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    size_t o3[TVL1_MAX_LEVELS];
    for (int s = 0; s < H->nlevels; s++) o3[s] = carve((size_t)H->lv[s].pitch * H->lv[s].h * sizeof(float));
    const size_t b0 = (size_t)H->lv[0].pitch * H->lv[0].h * sizeof(float);
    const size_t o31 = carve(b0), o32 = carve(b0);
    cudaError_t e = cudaMalloc(&H->garena, off);
    if (e != cudaSuccess) {
        H->garena = nullptr;
        return fail(TVL1_ERR_NOMEM, "cudaMalloc(%zu bytes) for the gamma planes failed: %s", off, cudaGetErrorString(e));
    }
    for (int s = 0; s < H->nlevels; s++) H->u3[s] = (float*)(H->garena + o3[s]);
    H->p31 = (float*)(H->garena + o31);
    H->p32 = (float*)(H->garena + o32);
    CK(cudaMemsetAsync(H->garena, 0, off, st));
    return TVL1_OK;
}

static int resolve_iterations(tvl1_handle* H)
{
    const tvl1_params& p = H->prm;
    if (p.inner_iterations > 0 && p.outer_iterations > 0) {
        H->inner = p.inner_iterations;
        H->outer = p.outer_iterations;
    } else {
        // the cv::cuda API's flat `iterations` = inner 30 x outer ceil(N/30)
        H->inner = p.inner_iterations > 0 ? p.inner_iterations : 30;
        const int n = p.iterations > 0 ? p.iterations : 300;
        H->outer = p.outer_iterations > 0 ? p.outer_iterations : (n + H->inner - 1) / H->inner;
    }
    return TVL1_OK;
}

static int check_params(const tvl1_params* p)
{
    if (!p) return fail(TVL1_ERR_INVALID, "params is null");
    if (p->nscales <= 0) return fail(TVL1_ERR_INVALID, "nscales must be > 0");   // CV_Assert(nscales > 0)
    if (p->nscales > TVL1_MAX_LEVELS) return fail(TVL1_ERR_INVALID, "nscales > %d", TVL1_MAX_LEVELS);
    if (p->warps <= 0 || p->warps > TVL1_MAX_WARPS) return fail(TVL1_ERR_INVALID, "warps out of range");
    if (!(p->scale_step > 0.0 && p->scale_step < 1.0)) return fail(TVL1_ERR_INVALID, "scaleStep must be in (0,1)");
    if (!(p->gamma == p->gamma)) return fail(TVL1_ERR_INVALID, "gamma is not a number");
    if (p->use_initial_flow) return fail(TVL1_ERR_UNSUPPORTED, "useInitialFlow is not supported (the reference never forwards it)");
    if (p->median_filtering != 1 && p->median_filtering != 3 && p->median_filtering != 5)   // cv::medianBlur's fp32 apertures
        return fail(TVL1_ERR_UNSUPPORTED, "medianFiltering must be 1 (off), 3 or 5");
    if (!(p->theta > 0.0)) return fail(TVL1_ERR_INVALID, "theta must be > 0");
    if (!(p->lambda >= 0.0)) return fail(TVL1_ERR_INVALID, "lambda must be >= 0");   // the threshold step assumes l_t * grad >= 0
    return TVL1_OK;
}

// the solve proper; everything is enqueued on `st`
static int calc_device(tvl1_handle* H, const uint8_t* f0, size_t pitch0, const uint8_t* f1, size_t pitch1,
                       int w, int h, float* d_u, float* d_v, size_t pitch_out, cudaStream_t st,
                       tvl1_stats* stats)
{
    if (!f0 || !f1 || !d_u || !d_v) return fail(TVL1_ERR_INVALID, "null image or flow pointer");
    if (w <= 0 || h <= 0) return fail(TVL1_ERR_INVALID, "empty image (%dx%d)", w, h);
    if (w > 32767 || h > 32767) return fail(TVL1_ERR_INVALID, "image side > 32767 (remap uses int16 coordinates)");
    if (pitch0 < (size_t)w || pitch1 < (size_t)w || pitch_out < (size_t)w * 4)
        return fail(TVL1_ERR_INVALID, "pitch smaller than a row");
    CK(cudaSetDevice(H->device));
    int rc;
    if ((rc = ensure_capacity(H, w, h, st))) return rc;
    if ((rc = ensure_gamma(H, st))) return rc;
    if ((rc = upload_cubic_table(H->device))) return rc;
    const tvl1_params& P = H->prm;
    const int L = H->nlevels, W = P.warps;
    const float l_t = (float)(P.lambda * P.theta);
    const float taut = (float)(P.tau / P.theta);
    const float theta = (float)P.theta;
    long long launches = 0;
    H->ev_used = 0;

    cudaEvent_t ev_begin, ev_pyr, ev_end;
    if ((rc = get_event(H, &ev_begin)) || (rc = get_event(H, &ev_pyr)) || (rc = get_event(H, &ev_end))) return rc;
    struct Span { cudaEvent_t a, b; int kind, level; };
    std::vector<Span> spans;
    // per-stage events only when the caller asked for them (tvl1_set_timing): thousands of
    // cudaEventRecord calls per pair otherwise
    const bool timed = stats != nullptr && H->timing;
    auto span_begin = [&](int kind, int level) -> int {
        if (!timed) return TVL1_OK;
        Span s; s.kind = kind; s.level = level;
        int r = get_event(H, &s.a); if (r) return r;
        r = get_event(H, &s.b); if (r) return r;
        cudaEventRecord(s.a, st);
        spans.push_back(s);
        return TVL1_OK;
    };
    auto span_end = [&]() { if (timed) cudaEventRecord(spans.back().b, st); };

    CK(cudaEventRecord(ev_begin, st));
    CK(cudaMemsetAsync(H->d_ctrl, 0, sizeof(Ctrl), st));
    // (1) pyramid: convert, then resize level by level
    if ((rc = launch_convert(f0, pitch0, w, h, H->lv[0].I0, H->lv[0].pitch, st))) return rc;
    if ((rc = launch_convert(f1, pitch1, w, h, H->lv[0].I1, H->lv[0].pitch, st))) return rc;
    launches += 2;
    for (int s = 1; s < L; s++) {
        const Level &a = H->lv[s - 1], &b = H->lv[s];
        if ((rc = launch_resize2(a.I0, a.I1, a.w, a.h, a.pitch, b.I0, b.I1, b.w, b.h, b.pitch, P.scale_step, 1.f, 0, st))) return rc;
        launches += 1;
    }
    {
        const Level& c = H->lv[L - 1];
        const size_t b = (size_t)c.pitch * c.h * sizeof(float);
        CK(cudaMemsetAsync(c.u1, 0, b, st));
        CK(cudaMemsetAsync(c.u2, 0, b, st));
        if (P.gamma != 0.0) CK(cudaMemsetAsync(H->u3[L - 1], 0, b, st));
    }
    CK(cudaEventRecord(ev_pyr, st));

    const size_t hdr = offsetof(Ctrl, iters);
    std::vector<int> counts((size_t)L * W, 0);   // iterations each (level, warp) needed
    double* dev_errlog = nullptr;                // developer aid: per-iteration error sums to stderr
    if (getenv("TVL1_DEV_ERRLOG")) cudaMalloc(&dev_errlog, sizeof(double) * L * W * H->inner * H->outer);
    struct ErrlogGuard { double*& p; ~ErrlogGuard() { if (p) cudaFree(p); p = nullptr; } } errlog_guard{dev_errlog};
    for (int s = L - 1; s >= 0; --s) {
        const Level& lv = H->lv[s];
        const size_t pb = (size_t)lv.pitch * lv.h * sizeof(float);
        const float scaled_eps = (float)(P.epsilon * P.epsilon * (double)(lv.w * lv.h));
        if ((rc = span_begin(3, s))) return rc;
        // p11 .. p22 start a level at zero: the level's first warp kernel writes the zeros (WarpArgs::pz)
        if (P.gamma != 0.0) { CK(cudaMemsetAsync(H->p31, 0, pb, st)); CK(cudaMemsetAsync(H->p32, 0, pb, st)); }
        span_end();

        IterArgs ia;
        ia.I1wx = H->I1wx; ia.I1wy = H->I1wy; ia.rho_c = H->rho;
        ia.u1[0] = lv.u1; ia.u1[1] = H->u1x; ia.u2[0] = lv.u2; ia.u2[1] = H->u2x;
        for (int k = 0; k < 2; k++) { ia.p11[k] = H->p[0][k]; ia.p12[k] = H->p[1][k]; ia.p21[k] = H->p[2][k]; ia.p22[k] = H->p[3][k]; }
        ia.w = lv.w; ia.h = lv.h; ia.pitch = lv.pitch;
        ia.l_t = l_t; ia.theta = theta; ia.taut = taut; ia.scaled_eps = scaled_eps; ia.one = 1.0f;
        ia.level = s; ia.ctrl = H->d_ctrl; ia.partials = H->d_partials; ia.errlog = nullptr;
        ia.mode = 0; ia.inner_max = H->inner;
        ia.tm = lv.tm_iter;
        const bool with_gamma = P.gamma != 0.0;
        GammaArgs ga;
        if (with_gamma) {
            ga.I1wx = H->I1wx; ga.I1wy = H->I1wy; ga.rho_c = H->rho;
            for (int k = 0; k < 2; k++) {
                ga.u1[k] = ia.u1[k]; ga.u2[k] = ia.u2[k];
                ga.p11[k] = ia.p11[k]; ga.p12[k] = ia.p12[k]; ga.p21[k] = ia.p21[k]; ga.p22[k] = ia.p22[k];
            }
            ga.u3 = H->u3[s]; ga.p31 = H->p31; ga.p32 = H->p32;
            ga.w = lv.w; ga.h = lv.h; ga.pitch = lv.pitch;
            ga.l_t = l_t; ga.theta = theta; ga.taut = taut; ga.gamma = (float)P.gamma; ga.scaled_eps = scaled_eps;
            ga.level = s; ga.slot = 0; ga.inner_max = H->inner; ga.mode = 1;
            ga.ctrl = H->d_ctrl; ga.partials = H->d_partials; ga.errlog = nullptr;
        }
        const bool fused = (long long)lv.w * lv.h >= H->fused_min_px && H->inner >= 2;
        const bool multi = !fused && H->multi_iter && H->inner >= 2;
        MedianArgs ma;
        ma.u1[0] = lv.u1; ma.u1[1] = H->u1x; ma.u2[0] = lv.u2; ma.u2[1] = H->u2x;
        ma.w = lv.w; ma.h = lv.h; ma.pitch = lv.pitch; ma.level = s; ma.ctrl = H->d_ctrl;
        memcpy(ma.tm, lv.tm_med, sizeof(ma.tm));
        WarpArgs wa;
        wa.I0 = lv.I0; wa.I1 = lv.I1;
        wa.u1[0] = lv.u1; wa.u1[1] = H->u1x; wa.u2[0] = lv.u2; wa.u2[1] = H->u2x;
        wa.I1w = nullptr; wa.I1wx = H->I1wx; wa.I1wy = H->I1wy; wa.grad = nullptr; wa.rho_c = H->rho;
        wa.w = lv.w; wa.h = lv.h; wa.pitch = lv.pitch; wa.level = s; wa.ctrl = H->d_ctrl; wa.one = 1.0f;
        memcpy(wa.tmI1, lv.tm_I1, sizeof(wa.tmI1));

        for (int wi = 0; wi < W; ++wi) {
            const int slot = s * W + wi;
            ia.slot = slot;
            if (dev_errlog) ia.errlog = dev_errlog + (size_t)slot * H->inner * H->outer;
            ma.slot = slot;
            // (2) warp I1 by the current flow; also re-arms the stop test (error = FLT_MAX)
            if ((rc = span_begin(1, s))) return rc;
            for (int k = 0; k < 4; k++) wa.pz[k] = wi == 0 ? H->p[k][0] : nullptr;
            if ((rc = launch_warp(wa, st))) return rc;
            launches++;
            span_end();
            // (3) outer iterations: median + up to `inner` primal-dual iterations.  Every kernel is a
            // no-op once the device-side stop flag is set, so iterations are enqueued in chunks without
            // knowing how many will really run; the first chunk is sized by the count the previous
            // warp (or, for the first warp, the coarser level) needed, and one 32-byte read-back per
            // chunk tells the host whether more are wanted.  The stop iteration is exactly the
            // oracle's; a wrong guess costs a few no-op launches or one extra host round trip.
            const int pred_total = wi > 0 ? counts[slot - 1] : (s < L - 1 ? counts[(s + 1) * W] : H->inner);
            int done_total = 0;
            for (int no = 0; no < H->outer; ++no) {
                if (P.median_filtering > 1) {
                    if ((rc = span_begin(2, s))) return rc;
                    if ((rc = P.median_filtering == 3 ? launch_median3(ma, 2, st) : launch_median(ma, 2, st))) return rc;
                    launches++;
                    span_end();
                }
                if ((rc = span_begin(0, s))) return rc;
                // a new outer iteration: no inner iteration done yet
                CK(cudaMemsetAsync(&H->d_ctrl->inner, 0, sizeof(int), st));
                int done_inner = 0, want = 0;
                for (;;) {
                    const int left_pred = pred_total - done_total - done_inner;
                    want = want == 0 ? (left_pred + 1 > 4 ? left_pred + 1 : 4) : 2 * want;
                    if (want > H->inner - done_inner) want = H->inner - done_inner;
                    if (with_gamma) {
                        // the three-channel iteration: two plain launches each, no-ops once the stop flag is set
                        ga.slot = slot;
                        ga.errlog = ia.errlog;
                        for (int k = 0; k < want; ++k)
                            if ((rc = launch_gamma_iteration(ga, st))) return rc;
                        launches += 2 * want;
                    } else if (fused && H->coop_outer) {
                        // the whole inner loop of this outer iteration in one cooperative launch
                        IterArgs io = ia;
                        io.mode = 2;
                        if ((rc = launch_outer(io, st))) return rc;
                        launches += 1;
                    } else if (fused) {
                        // temporally blocked schedule: [fused pair | single] slot pairs.  Which slot
                        // of a pair really runs is decided on the device: the pair while the stop is
                        // not imminent and two iterations still fit, the single after an overshoot
                        // (replay), near the stop, or for the last iteration.
                        IterArgs i2 = ia, i1 = ia;
                        i2.mode = 2; i1.mode = 1;
                        const int groups = (want + 1) / 2;
                        for (int k = 0; k < groups; ++k) {
                            if ((rc = launch_iterate2(i2, st))) return rc;
                            if ((rc = launch_iterate(i1, st))) return rc;
                        }
                        launches += 2 * groups;
                    } else if (multi) {
                        // small level: one cooperative launch runs the rest of this outer iteration
                        IterArgs im = ia;
                        im.mode = 3;
                        if ((rc = launch_iterate_multi(im, st))) return rc;
                        launches += 1;
                    } else {
                        IterArgs i3 = ia;
                        i3.mode = 3;
                        for (int k = 0; k < want; ++k)
                            if ((rc = launch_iterate(i3, st))) return rc;
                        launches += want;
                    }
                    if ((rc = copy_words(H->d_ctrl, H->h_ctrl, (int)(hdr / 4), st))) return rc;   // no copy engine: see copy_words
                    CK(cudaStreamSynchronize(st));
                    done_inner = H->h_ctrl->inner;
                    if (H->h_ctrl->done || done_inner >= H->inner) break;
                }
                span_end();
                done_total += done_inner;
                if (H->h_ctrl->done) break;
            }
            counts[slot] = done_total;
            if (dev_errlog && done_total > 0) {
                std::vector<double> e(done_total);
                cudaMemcpy(e.data(), ia.errlog, sizeof(double) * done_total, cudaMemcpyDeviceToHost);
                fprintf(stderr, "errlog L%d w%d n=%d e/eps:", s, wi, done_total);
                for (int k = 0; k < done_total; k++) fprintf(stderr, " %.3g", e[k] / scaled_eps);
                fprintf(stderr, "\n");
            }
        }
        if (s == 0) break;
        // flow upsample to the next finer level (A.2): resize to its size, times 1/scaleStep
        const Level& up = H->lv[s - 1];
        const int uc = H->h_ctrl->ucur[s];
        const float mul = (float)(1 / P.scale_step);
        if ((rc = span_begin(3, s))) return rc;
        if ((rc = launch_resize2(uc ? H->u1x : lv.u1, uc ? H->u2x : lv.u2, lv.w, lv.h, lv.pitch, up.u1, up.u2, up.w, up.h,
                                 up.pitch, 0.0, mul, 1, st))) return rc;
        launches += 1;
        if (with_gamma) {   // u3 is zoomed like the flow but not scaled
            if ((rc = launch_resize(H->u3[s], lv.w, lv.h, lv.pitch, H->u3[s - 1], up.w, up.h, up.pitch, 0.0, 1.f, 0, st))) return rc;
            launches += 1;
        }
        span_end();
    }
    // A.8: planar output
    {
        const int uc = H->h_ctrl->ucur[0];
        const Level& l0 = H->lv[0];
        CK(cudaMemcpy2DAsync(d_u, pitch_out, uc ? H->u1x : l0.u1, (size_t)l0.pitch * 4, (size_t)w * 4, h, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpy2DAsync(d_v, pitch_out, uc ? H->u2x : l0.u2, (size_t)l0.pitch * 4, (size_t)w * 4, h, cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaEventRecord(ev_end, st));
    if ((rc = copy_words(H->d_ctrl, H->h_ctrl, (int)(sizeof(Ctrl) / 4), st))) return rc;
    CK(cudaStreamSynchronize(st));

    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->levels = L;
        stats->warps = W;
        double bytes = 70.0 * (double)w * h;
        for (int s = 0; s < L; s++) {
            stats->width[s] = H->lv[s].w;
            stats->height[s] = H->lv[s].h;
            const double px = (double)H->lv[s].w * H->lv[s].h;
            double per_px = 12.0;
            for (int wi = 0; wi < W; wi++) {
                const int n = H->h_ctrl->iters[s * W + wi], o = H->h_ctrl->outer[s * W + wi];
                stats->iters[s * W + wi] = n;
                stats->outer[s * W + wi] = o;
                stats->total_iterations += n;
                stats->px_iterations += (long long)px * n;
                per_px += 40.0 + 64.0 * n + 16.0 * o;
            }
            bytes += px * per_px;
        }
        stats->algorithmic_bytes = bytes;
        stats->launches = launches;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev_begin, ev_end);
        stats->ms_total = ms;
        cudaEventElapsedTime(&ms, ev_begin, ev_pyr);
        stats->ms_pyramid = ms;
        for (const Span& sp : spans) {
            cudaEventElapsedTime(&ms, sp.a, sp.b);
            switch (sp.kind) {
                case 0: stats->ms_iterate += ms; stats->ms_iterate_level[sp.level] += ms; break;
                case 1: stats->ms_warp += ms; break;
                case 2: stats->ms_median += ms; break;
                default: stats->ms_other += ms; break;
            }
        }
    }
    return TVL1_OK;
}

}  // namespace tvl1

// =================================================================== C ABI

extern "C" {

const char* tvl1_version(void) { return "fibsem-optflow_b200 0.1 (sm_100a)"; }
const char* tvl1_last_error(void) { return tvl1::g_err; }

void tvl1_default_params(tvl1_params* p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    // generate_TV_args defaults, reference src/optflow.cpp:503-512
    p->tau = 0.25; p->lambda = 0.05; p->theta = 0.3; p->nscales = 10; p->warps = 5;
    p->epsilon = 0.01; p->iterations = 300; p->scale_step = 0.8; p->gamma = 0.0;
    p->use_initial_flow = 0;
    p->inner_iterations = 0; p->outer_iterations = 0;   // derived from iterations
    p->median_filtering = 5;                             // CPU-class semantics
}

int tvl1_create(const tvl1_params* p, int device, tvl1_handle** out)
{
    if (!out) return fail(TVL1_ERR_INVALID, "out is null");
    *out = nullptr;
    int rc = check_params(p);
    if (rc) return rc;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(TVL1_ERR_CUDA, "no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(TVL1_ERR_INVALID, "device %d out of range (%d devices)", device, n);
    CK(cudaSetDevice(device));
    tvl1_handle* H = new tvl1_handle();
    H->device = device;
    H->prm = *p;
    resolve_iterations(H);
    e = cudaMallocHost(&H->h_ctrl, sizeof(Ctrl));
    if (e != cudaSuccess) { delete H; return fail(TVL1_ERR_CUDA, "cudaMallocHost: %s", cudaGetErrorString(e)); }
    memset(H->h_ctrl, 0, sizeof(Ctrl));
    e = cudaStreamCreateWithFlags(&H->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { cudaFreeHost(H->h_ctrl); delete H; return fail(TVL1_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    rc = upload_cubic_table(device);
    if (rc) { tvl1_destroy(H); return rc; }
    *out = H;
    return TVL1_OK;
}

void tvl1_destroy(tvl1_handle* H)
{
    if (!H) return;
    cudaSetDevice(H->device);
    tvl1::release_arena(H);
    if (H->h_ctrl) cudaFreeHost(H->h_ctrl);
    if (H->d_f0) cudaFree(H->d_f0);
    if (H->d_f1) cudaFree(H->d_f1);
    if (H->d_uo) cudaFree(H->d_uo);
    if (H->d_vo) cudaFree(H->d_vo);
    tvl1::sampler_release(H->samp);
    tvl1::features_release(H->feat);
    for (int i = 0; i < TVL1_STACK_SLOTS; i++) { if (H->st_slice[i]) cudaFree(H->st_slice[i]); if (H->st_up[i]) cudaEventDestroy(H->st_up[i]); }
    for (int i = 0; i < 2; i++) if (H->st_raw[i]) cudaFree(H->st_raw[i]);
    for (int i = 0; i < 2; i++) {
        for (int j = 0; j < 2; j++) if (H->st_flow[i][j]) cudaFree(H->st_flow[i][j]);
        if (H->st_ready[i]) cudaEventDestroy(H->st_ready[i]);
        if (H->st_down[i]) cudaEventDestroy(H->st_down[i]);
    }
    if (H->st_in) cudaStreamDestroy(H->st_in);
    if (H->st_out) cudaStreamDestroy(H->st_out);
    for (cudaEvent_t e : H->events) cudaEventDestroy(e);
    if (H->own_stream) cudaStreamDestroy(H->own_stream);
    delete H;
}

int tvl1_set_params(tvl1_handle* H, const tvl1_params* p)
{
    if (!H) return fail(TVL1_ERR_INVALID, "handle is null");
    int rc = check_params(p);
    if (rc) return rc;
    H->prm = *p;
    return resolve_iterations(H);
}

int tvl1_set_option(tvl1_handle* H, const char* key, double value)
{
    if (!H || !key) return fail(TVL1_ERR_INVALID, "null handle or key");
    if (!strcmp(key, "fused_min_px")) { H->fused_min_px = value < 0 ? 0 : (long long)value; return TVL1_OK; }
    if (!strcmp(key, "multi_iter")) { H->multi_iter = value != 0; return TVL1_OK; }
    if (!strcmp(key, "coop_outer")) { H->coop_outer = value != 0; return TVL1_OK; }
    return fail(TVL1_ERR_INVALID, "unknown option '%s'", key);
}

int tvl1_set_timing(tvl1_handle* H, int enabled)
{
    if (!H) return fail(TVL1_ERR_INVALID, "handle is null");
    H->timing = enabled != 0;
    return TVL1_OK;
}

int tvl1_calc_u8(tvl1_handle* H, const uint8_t* d_frame0, size_t pitch0, const uint8_t* d_frame1, size_t pitch1,
                 int width, int height, float* d_u, float* d_v, size_t pitch_out, void* stream, tvl1_stats* stats)
{
    if (!H) return fail(TVL1_ERR_INVALID, "handle is null");
    return calc_device(H, d_frame0, pitch0, d_frame1, pitch1, width, height, d_u, d_v, pitch_out,
                       (cudaStream_t)stream, stats);
}

int tvl1_calc_u8_host(tvl1_handle* H, const uint8_t* h_frame0, size_t pitch0, const uint8_t* h_frame1, size_t pitch1,
                      int width, int height, float* h_u, float* h_v, size_t pitch_out, tvl1_stats* stats)
{
    if (!H) return fail(TVL1_ERR_INVALID, "handle is null");
    if (!h_frame0 || !h_frame1 || !h_u || !h_v) return fail(TVL1_ERR_INVALID, "null image or flow pointer");
    if (width <= 0 || height <= 0) return fail(TVL1_ERR_INVALID, "empty image (%dx%d)", width, height);
    if (pitch0 < (size_t)width || pitch1 < (size_t)width || pitch_out < (size_t)width * 4)
        return fail(TVL1_ERR_INVALID, "pitch smaller than a row");
    CK(cudaSetDevice(H->device));
    if (H->stage_w != width || H->stage_h != height) {
        if (H->d_f0) cudaFree(H->d_f0);
        if (H->d_f1) cudaFree(H->d_f1);
        if (H->d_uo) cudaFree(H->d_uo);
        if (H->d_vo) cudaFree(H->d_vo);
        H->d_f0 = H->d_f1 = nullptr; H->d_uo = H->d_vo = nullptr; H->stage_w = H->stage_h = 0;
        H->stage_pitch8 = (size_t)round_up(width, 128);
        CK(cudaMalloc(&H->d_f0, H->stage_pitch8 * height));
        CK(cudaMalloc(&H->d_f1, H->stage_pitch8 * height));
        CK(cudaMalloc(&H->d_uo, (size_t)width * height * 4));
        CK(cudaMalloc(&H->d_vo, (size_t)width * height * 4));
        H->stage_w = width; H->stage_h = height;
    }
    cudaStream_t st = H->own_stream;
    CK(cudaMemcpy2DAsync(H->d_f0, H->stage_pitch8, h_frame0, pitch0, width, height, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpy2DAsync(H->d_f1, H->stage_pitch8, h_frame1, pitch1, width, height, cudaMemcpyHostToDevice, st));
    int rc = calc_device(H, H->d_f0, H->stage_pitch8, H->d_f1, H->stage_pitch8, width, height, H->d_uo, H->d_vo,
                         (size_t)width * 4, st, stats);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(h_u, pitch_out, H->d_uo, (size_t)width * 4, (size_t)width * 4, height, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpy2DAsync(h_v, pitch_out, H->d_vo, (size_t)width * 4, (size_t)width * 4, height, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return TVL1_OK;
}

static int launch_mask_flow(const uint8_t* f1, size_t pitch1, int w, int h, float* u, float* v, size_t pitch_f, int add_grid,
                            cudaStream_t st)
{
    dim3 b(32, 8);
    dim3 g(cdiv(cdiv(w, 4), 32), cdiv(h, 8));
    k_mask_flow<<<g, b, 0, st>>>(f1, pitch1, w, h, u, v, pitch_f, add_grid);
    CK(cudaGetLastError());
    return TVL1_OK;
}

int tvl1_finish_flow_u8(tvl1_handle* H, const uint8_t* d_frame1, size_t pitch1, int width, int height,
                        float* d_u, float* d_v, size_t pitch_out, int add_grid, void* stream)
{
    if (!H) return fail(TVL1_ERR_INVALID, "handle is null");
    if (!d_u || !d_v || width <= 0 || height <= 0 || pitch_out % 4) return fail(TVL1_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(H->device));
    return launch_mask_flow(d_frame1, pitch1, width, height, d_u, d_v, pitch_out / 4, add_grid > 0 ? 1 : (add_grid < 0 ? -1 : 0),
                            (cudaStream_t)stream);
}

int tvl1_mask_flow_u8(tvl1_handle* H, const uint8_t* d_frame1, size_t pitch1, int width, int height,
                      float* d_u, float* d_v, size_t pitch_out, void* stream)
{
    if (!d_frame1) return fail(TVL1_ERR_INVALID, "frame1 is null");
    return tvl1_finish_flow_u8(H, d_frame1, pitch1, width, height, d_u, d_v, pitch_out, 0, stream);
}

// ---- stack of adjacent slices

static int stack_reserve(tvl1_handle* H, int w, int h)
{
    if (!H->st_in) {
        CK(cudaStreamCreateWithFlags(&H->st_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&H->st_out, cudaStreamNonBlocking));
        for (int i = 0; i < TVL1_STACK_SLOTS; i++) CK(cudaEventCreateWithFlags(&H->st_up[i], cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) {
            CK(cudaEventCreateWithFlags(&H->st_ready[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&H->st_down[i], cudaEventDisableTiming));
        }
    }
    if (H->st_w == w && H->st_h == h) return TVL1_OK;
    for (int i = 0; i < TVL1_STACK_SLOTS; i++) { if (H->st_slice[i]) cudaFree(H->st_slice[i]); H->st_slice[i] = nullptr; }
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) { if (H->st_flow[i][j]) cudaFree(H->st_flow[i][j]); H->st_flow[i][j] = nullptr; }
    H->st_w = H->st_h = 0;
    H->st_pitch8 = (size_t)round_up(w, 128);
    for (int i = 0; i < TVL1_STACK_SLOTS; i++) CK(cudaMalloc(&H->st_slice[i], H->st_pitch8 * h));
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) CK(cudaMalloc(&H->st_flow[i][j], (size_t)w * h * sizeof(float)));
    H->st_w = w; H->st_h = h;
    return TVL1_OK;
}

int tvl1_stack_run(tvl1_handle* H, const tvl1_stack_io* io, float* ms_total)
{
    if (!H || !io) return fail(TVL1_ERR_INVALID, "null handle or io");
    const int n = io->n_slices;
    const int rw = io->width, rh = io->height;          // size of the slices as given
    int w = rw, h = rh;                                 // size the solver works at
    const bool pre = io->prescale > 0.0 && io->prescale != 1.0;
    const bool listed = io->pair_p != nullptr || io->pair_q != nullptr;   // explicit (p, q) pairs instead of (k, k+1)
    if (!io->h_slices || n < 2 || rw <= 0 || rh <= 0 || io->pitch < (size_t)rw)
        return fail(TVL1_ERR_INVALID, "a stack needs >= 2 slices of non-zero size");
    if (listed && (!io->pair_p || !io->pair_q || io->n_pairs < 1)) return fail(TVL1_ERR_INVALID, "pair_p, pair_q and n_pairs go together");
    const int npairs = listed ? io->n_pairs : n - 1;
    if (pre) { int r = tvl1_prescaled_size(rw, rh, io->prescale, &w, &h); if (r) return r; }
    if ((io->h_u == nullptr) != (io->h_v == nullptr)) return fail(TVL1_ERR_INVALID, "h_u and h_v go together");
    if (io->h_u && io->pitch_out < (size_t)w * 4) return fail(TVL1_ERR_INVALID, "pitch_out smaller than a row");
    if (io->npoints >= 0 && (!io->px || !io->py || !io->qx || !io->qy || !io->w || !io->n_out))
        return fail(TVL1_ERR_INVALID, "match output arrays missing");
    for (int k = 0; k < n; k++) if (!io->h_slices[k]) return fail(TVL1_ERR_INVALID, "slice %d is null", k);
    auto pp = [&](int k) { return listed ? io->pair_p[k] : k; };
    auto pq = [&](int k) { return listed ? io->pair_q[k] : k + 1; };
    for (int k = 0; k < npairs; k++)
        if (pp(k) < 0 || pp(k) >= n || pq(k) < 0 || pq(k) >= n) return fail(TVL1_ERR_INVALID, "pair %d names a slice outside the stack", k);
    CK(cudaSetDevice(H->device));
    int rc = stack_reserve(H, w, h);
    if (rc) return rc;
    if (pre && H->st_raw_bytes < (size_t)rw * rh) {
        for (int i = 0; i < 2; i++) {
            if (H->st_raw[i]) cudaFree(H->st_raw[i]);
            H->st_raw[i] = nullptr;
        }
        H->st_raw_bytes = 0;
        for (int i = 0; i < 2; i++) CK(cudaMalloc(&H->st_raw[i], (size_t)rw * rh));
        H->st_raw_bytes = (size_t)rw * rh;
    }
    cudaStream_t cs = H->own_stream;
    const size_t p8 = H->st_pitch8;
    const int cap = io->npoints > 0 ? io->npoints : 1;
    struct EventPair {
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } tev;
    CK(cudaEventCreate(&tev.a));
    CK(cudaEventCreate(&tev.b));
    cudaEvent_t t0 = tev.a, t1 = tev.b;
    CK(cudaEventRecord(t0, cs));
    // Device slots for the (prescaled) slices, kept by slice index: the slice two adjacent pairs share goes
    // up once (src/optflow.cpp:97-103), and the frames of pair k+1 arrive on the copy stream while pair k is
    // being solved.  `held[s]`: last pair that reads slot s -- a slot is only refilled once that pair's solve
    // has completed (the solves are synchronous to the host, so "completed" is known here).
    const int NS = TVL1_STACK_SLOTS;
    int tag[NS], held[NS];
    for (int s = 0; s < NS; s++) { tag[s] = -1; held[s] = -1; }
    int uploads = 0;
    auto find = [&](int slice) { for (int s = 0; s < NS; s++) if (tag[s] == slice) return s; return -1; };
    // slot for `slice` if it is resident or can be brought in without touching a slot that pair `busy` (or a later
    // one) reads
    auto stage = [&](int slice, int for_pair, int busy, int* out) -> int {
        int s = find(slice);
        if (s < 0) {
            for (int c = 0; c < NS; c++)
                if (held[c] < busy && (s < 0 || held[c] < held[s])) s = c;
            if (s < 0) { *out = -1; return TVL1_OK; }   // every slot is still in use: later
            tag[s] = slice;
            if (!pre) {
                CK(cudaMemcpy2DAsync(H->st_slice[s], p8, io->h_slices[slice], io->pitch, w, h, cudaMemcpyHostToDevice, H->st_in));
            } else {
                // raw slice -> device, then the loader's 8-bit resize (src/optflow.cpp:111,124) on the
                // copy stream: the solver never sees the raw size.  Two staging buffers, re-used in stream
                // order (an upload follows the prescale that last read the buffer, on the same stream)
                uint8_t* raw = H->st_raw[uploads & 1];
                CK(cudaMemcpy2DAsync(raw, (size_t)rw, io->h_slices[slice], io->pitch, rw, rh, cudaMemcpyHostToDevice, H->st_in));
                int r = tvl1_prescale_u8(raw, (size_t)rw, rw, rh, io->prescale, H->st_slice[s], p8, H->st_in);
                if (r) return r;
            }
            uploads++;
            CK(cudaEventRecord(H->st_up[s], H->st_in));
        }
        if (held[s] < for_pair) held[s] = for_pair;
        *out = s;
        return TVL1_OK;
    };
    long long rand_skip = 0;
    for (int k = 0; k < npairs; k++) {
        const int fb = k & 1;
        int s0 = -1, s1 = -1, tmp = -1;
        // this pair's frames (pair k-1 has completed: only pair k itself pins slots now)
        if ((rc = stage(pp(k), k, k, &s0)) || (rc = stage(pq(k), k, k, &s1))) return rc;
        if (s0 < 0 || s1 < 0) return fail(TVL1_ERR_CUDA, "no free slice slot (internal)");
        // the next pair's frames go up now, into slots this pair does not read
        if (k + 1 < npairs) {
            if ((rc = stage(pp(k + 1), k + 1, k, &tmp)) || (rc = stage(pq(k + 1), k + 1, k, &tmp))) return rc;
        }
        CK(cudaStreamWaitEvent(cs, H->st_up[s0], 0));
        CK(cudaStreamWaitEvent(cs, H->st_up[s1], 0));
        if (k >= 2 && io->h_u) CK(cudaStreamWaitEvent(cs, H->st_down[fb], 0));   // flow buffer still draining
        float* du = H->st_flow[fb][0];
        float* dv = H->st_flow[fb][1];
        rc = calc_device(H, H->st_slice[s0], p8, H->st_slice[s1], p8, w, h, du, dv, (size_t)w * 4, cs,
                         io->stats ? &io->stats[k] : nullptr);
        if (rc) return rc;
        if (io->apply_mask && (rc = launch_mask_flow(H->st_slice[s1], p8, w, h, du, dv, (size_t)w, 0, cs))) return rc;
        if (io->npoints >= 0) {
            long long used = 0;
            const size_t o = (size_t)k * cap;
            rc = tvl1_sample_matches_skip(H, H->st_slice[s0], p8, H->st_slice[s1], p8, du, dv, (size_t)w * 4, w, h,
                                          0, 0, 0, 0, io->scale, io->npoints, io->seed, io->seed < 0 ? rand_skip : 0,
                                          io->px + o, io->py + o, io->qx + o, io->qy + o, io->w + o, nullptr,
                                          &io->n_out[k], &used, cs);
            if (rc) return rc;
            rand_skip += used;
        } else if (io->apply_mask) {
            CK(cudaStreamSynchronize(cs));   // slot s1 is read by the mask kernel: done before it can be refilled
        }
        if (io->h_u) {
            CK(cudaEventRecord(H->st_ready[fb], cs));
            CK(cudaStreamWaitEvent(H->st_out, H->st_ready[fb], 0));
            CK(cudaMemcpy2DAsync(io->h_u[k], io->pitch_out, du, (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyDeviceToHost, H->st_out));
            CK(cudaMemcpy2DAsync(io->h_v[k], io->pitch_out, dv, (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyDeviceToHost, H->st_out));
            CK(cudaEventRecord(H->st_down[fb], H->st_out));
        }
    }
    CK(cudaStreamSynchronize(H->st_out));
    CK(cudaStreamSynchronize(H->st_in));
    CK(cudaEventRecord(t1, cs));
    CK(cudaEventSynchronize(t1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    if (ms_total) *ms_total = ms;
    return TVL1_OK;
}

// ---- stage-level entry points

// CUDA-event bracket around the launches of the most recent stage-level call (tvl1_k_last_ms)
static thread_local cudaEvent_t t_ev0 = nullptr, t_ev1 = nullptr;
static void stage_begin(cudaStream_t st)
{
    if (!t_ev0) { cudaEventCreate(&t_ev0); cudaEventCreate(&t_ev1); }
    cudaEventRecord(t_ev0, st);
}
static void stage_end(cudaStream_t st) { cudaEventRecord(t_ev1, st); }

int tvl1_k_last_ms(float* ms)
{
    if (!ms || !t_ev0) return fail(TVL1_ERR_INVALID, "no stage-level call has been timed on this thread");
    CK(cudaEventSynchronize(t_ev1));
    CK(cudaEventElapsedTime(ms, t_ev0, t_ev1));
    return TVL1_OK;
}

int tvl1_k_convert_u8(const uint8_t* d_src, size_t pitch_bytes, int w, int h, float* d_dst, int pitch, void* stream)
{
    if (!d_src || !d_dst || w <= 0 || h <= 0 || pitch % 4 || pitch < w || pitch_bytes < (size_t)w)
        return fail(TVL1_ERR_INVALID, "bad argument");
    return launch_convert(d_src, pitch_bytes, w, h, d_dst, pitch, (cudaStream_t)stream);
}

int tvl1_k_resize(const float* d_src, int sw, int sh, int spitch, float* d_dst, int dw, int dh, int dpitch,
                  double inv_scale, float mul, void* stream)
{
    if (!d_src || !d_dst || sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) return fail(TVL1_ERR_INVALID, "bad argument");
    return launch_resize(d_src, sw, sh, spitch, d_dst, dw, dh, dpitch, inv_scale, mul, mul != 1.f, (cudaStream_t)stream);
}

int tvl1_k_centered_gradient(const float* d_src, int w, int h, int pitch, float* d_dx, float* d_dy, void* stream)
{
    if (!d_src || !d_dx || !d_dy || w <= 0 || h <= 0) return fail(TVL1_ERR_INVALID, "bad argument");
    return launch_gradient(d_src, w, h, pitch, d_dx, d_dy, (cudaStream_t)stream);
}

int tvl1_k_warp(const float* d_I0, const float* d_I1, const float* d_u1, const float* d_u2, int w, int h,
                int pitch, float* d_I1w, float* d_I1wx, float* d_I1wy, float* d_grad, float* d_rho_c, void* stream)
{
    if (!d_I0 || !d_I1 || !d_u1 || !d_u2 || !d_I1wx || !d_I1wy || !d_rho_c || w <= 0 || h <= 0 || pitch % 4 || pitch < w)
        return fail(TVL1_ERR_INVALID, "bad argument");
    int dev = 0;
    CK(cudaGetDevice(&dev));
    int rc = upload_cubic_table(dev);
    if (rc) return rc;
    WarpArgs a;
    a.I0 = d_I0; a.I1 = d_I1;
    a.u1[0] = a.u1[1] = d_u1; a.u2[0] = a.u2[1] = d_u2;
    a.I1w = d_I1w; a.I1wx = d_I1wx; a.I1wy = d_I1wy; a.grad = d_grad; a.rho_c = d_rho_c;
    a.w = w; a.h = h; a.pitch = pitch; a.level = -1; a.ctrl = nullptr; a.one = 1.0f;
    for (int k = 0; k < 4; k++) a.pz[k] = nullptr;
    for (int k = 0; k < 2; k++)
        if ((rc = make_plane_map(&a.tmI1[k], d_I1, pitch, h, TVL1_WP_RW, k == 0 ? 16 : TVL1_WP_RH))) return rc;
    stage_begin((cudaStream_t)stream);
    rc = launch_warp(a, (cudaStream_t)stream);
    stage_end((cudaStream_t)stream);
    return rc;
}

static int k_iterate_impl(int fused, const float* d_I1wx, const float* d_I1wy, const float* d_grad, const float* d_rho_c,
                   float* d_u1, float* d_u2, float* d_p11, float* d_p12, float* d_p21, float* d_p22,
                   int w, int h, int pitch, float l_t, float theta, float taut, int n, double* errors, void* stream)
{
    if (!d_I1wx || !d_I1wy || !d_grad || !d_rho_c || !d_u1 || !d_u2 || !d_p11 || !d_p12 || !d_p21 || !d_p22 ||
        w <= 0 || h <= 0 || pitch % 4 || pitch < w || n < 0)
        return fail(TVL1_ERR_INVALID, "bad argument");
    if (n > TVL1_MAX_LEVELS * TVL1_MAX_WARPS) return fail(TVL1_ERR_INVALID, "n too large");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t pb = (size_t)pitch * h * sizeof(float);
    const size_t nb = iterate_max_blocks();
    char* tmp = nullptr;
    // the two-iteration pass fills its ring by bulk tensor copies of whole sections (IterMaps): its operands are
    // staged as 15 equally spaced planes {I1wx, I1wy, rho_c | u1, u2 | u1', u2' | p11..p22 | p11'..p22'}; the
    // one-iteration kernel works on the caller's planes with 6 twins
    const int nplanes = fused ? 15 : 6;
    const size_t ctrl_off = nplanes * pb, part_off = ctrl_off + 1024 * ((sizeof(Ctrl) + 1023) / 1024);
    const size_t log_off = part_off + nb * sizeof(double);
    const size_t total = log_off + (size_t)(n + 1) * sizeof(double);
    CK(cudaMalloc(&tmp, total));
    cudaError_t e = cudaMemsetAsync(tmp, 0, total, st);
    if (e != cudaSuccess) { cudaFree(tmp); return fail(TVL1_ERR_CUDA, "memset: %s", cudaGetErrorString(e)); }
    float* tw[15];
    for (int k = 0; k < nplanes; k++) tw[k] = (float*)(tmp + k * pb);
    float* const user[6] = {d_u1, d_u2, d_p11, d_p12, d_p21, d_p22};
    IterArgs a;
    if (fused) {
        const float* cs[3] = {d_I1wx, d_I1wy, d_rho_c};
        for (int k = 0; k < 3; k++) cudaMemcpyAsync(tw[k], cs[k], pb, cudaMemcpyDeviceToDevice, st);
        for (int k = 0; k < 2; k++) cudaMemcpyAsync(tw[3 + k], user[k], pb, cudaMemcpyDeviceToDevice, st);
        for (int k = 0; k < 4; k++) cudaMemcpyAsync(tw[7 + k], user[2 + k], pb, cudaMemcpyDeviceToDevice, st);
        a.I1wx = tw[0]; a.I1wy = tw[1]; a.rho_c = tw[2];
        a.u1[0] = tw[3]; a.u2[0] = tw[4]; a.u1[1] = tw[5]; a.u2[1] = tw[6];
        a.p11[0] = tw[7]; a.p12[0] = tw[8]; a.p21[0] = tw[9]; a.p22[0] = tw[10];
        a.p11[1] = tw[11]; a.p12[1] = tw[12]; a.p21[1] = tw[13]; a.p22[1] = tw[14];
    } else {
        a.I1wx = d_I1wx; a.I1wy = d_I1wy; a.rho_c = d_rho_c;   // grad is recomputed in the kernel
        a.u1[0] = d_u1; a.u1[1] = tw[0]; a.u2[0] = d_u2; a.u2[1] = tw[1];
        a.p11[0] = d_p11; a.p11[1] = tw[2]; a.p12[0] = d_p12; a.p12[1] = tw[3];
        a.p21[0] = d_p21; a.p21[1] = tw[4]; a.p22[0] = d_p22; a.p22[1] = tw[5];
    }
    a.w = w; a.h = h; a.pitch = pitch; a.l_t = l_t; a.theta = theta; a.taut = taut; a.one = 1.0f;
    a.scaled_eps = -1.f;   // never stops
    a.level = 0; a.slot = 0; a.mode = 0; a.inner_max = 1 << 30;
    a.ctrl = (Ctrl*)(tmp + ctrl_off);
    a.partials = (double*)(tmp + part_off);
    a.errlog = (double*)(tmp + log_off);
    int rc = TVL1_OK;
    if (fused && (rc = make_iter_maps(a))) { cudaFree(tmp); return rc; }
    stage_begin(st);
    if (fused == 2) {
        // the shipped schedule: ONE cooperative k_outer launch runs all n iterations (two-iteration
        // passes, plus one single pass when n is odd); scaled_eps < 0 keeps the stop test from firing
        a.mode = 2; a.inner_max = n;
        if (n > 0) rc = launch_outer(a, st);
    } else if (fused) { for (int i = 0; i + 1 < n && !rc; i += 2) rc = launch_iterate2(a, st); }
    else { for (int i = 0; i < n && !rc; i++) rc = launch_iterate(a, st); }
    stage_end(st);
    // passes made = buffer flips: n single passes, n/2 fused ones, n/2 + n%2 inside k_outer
    const int flips = fused == 2 ? n / 2 + (n & 1) : (fused ? n / 2 : n);
    if (!rc && (fused || (flips & 1))) {
        const int t = flips & 1;
        float* const src[6] = {a.u1[t], a.u2[t], a.p11[t], a.p12[t], a.p21[t], a.p22[t]};
        for (int k = 0; k < 6 && !rc; k++)
            if (cudaMemcpyAsync(user[k], src[k], pb, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
                rc = fail(TVL1_ERR_CUDA, "copy back failed");
    }
    if (!rc && errors && n > 0 &&
        (e = cudaMemcpyAsync(errors, tmp + log_off, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess)
        rc = fail(TVL1_ERR_CUDA, "error log copy failed: %s", cudaGetErrorString(e));
    e = cudaStreamSynchronize(st);
    if (!rc && e != cudaSuccess) rc = fail(TVL1_ERR_CUDA, "iterate: %s", cudaGetErrorString(e));
    cudaFree(tmp);
    return rc;
}

int tvl1_k_iterate(const float* d_I1wx, const float* d_I1wy, const float* d_grad, const float* d_rho_c,
                   float* d_u1, float* d_u2, float* d_p11, float* d_p12, float* d_p21, float* d_p22,
                   int w, int h, int pitch, float l_t, float theta, float taut, int n, double* errors, void* stream)
{
    return k_iterate_impl(0, d_I1wx, d_I1wy, d_grad, d_rho_c, d_u1, d_u2, d_p11, d_p12, d_p21, d_p22, w, h, pitch,
                          l_t, theta, taut, n, errors, stream);
}

int tvl1_k_iterate_fused2(const float* d_I1wx, const float* d_I1wy, const float* d_grad, const float* d_rho_c,
                          float* d_u1, float* d_u2, float* d_p11, float* d_p12, float* d_p21, float* d_p22,
                          int w, int h, int pitch, float l_t, float theta, float taut, int n, double* errors,
                          void* stream)
{
    if (n & 1) return fail(TVL1_ERR_INVALID, "the fused kernel advances two iterations per launch: n must be even");
    return k_iterate_impl(1, d_I1wx, d_I1wy, d_grad, d_rho_c, d_u1, d_u2, d_p11, d_p12, d_p21, d_p22, w, h, pitch,
                          l_t, theta, taut, n, errors, stream);
}

int tvl1_k_outer(const float* d_I1wx, const float* d_I1wy, const float* d_grad, const float* d_rho_c,
                 float* d_u1, float* d_u2, float* d_p11, float* d_p12, float* d_p21, float* d_p22,
                 int w, int h, int pitch, float l_t, float theta, float taut, int n, double* errors, void* stream)
{
    return k_iterate_impl(2, d_I1wx, d_I1wy, d_grad, d_rho_c, d_u1, d_u2, d_p11, d_p12, d_p21, d_p22, w, h, pitch,
                          l_t, theta, taut, n, errors, stream);
}

int tvl1_k_iterate_gamma(const float* d_I1wx, const float* d_I1wy, const float* d_rho_c,
                         float* d_u1, float* d_u2, float* d_u3, float* d_p11, float* d_p12,
                         float* d_p21, float* d_p22, float* d_p31, float* d_p32, int w, int h,
                         int pitch, float l_t, float theta, float taut, float gamma, int n,
                         double* errors, void* stream)
{
    if (!d_I1wx || !d_I1wy || !d_rho_c || !d_u1 || !d_u2 || !d_u3 || !d_p11 || !d_p12 || !d_p21 || !d_p22 ||
        !d_p31 || !d_p32 || w <= 0 || h <= 0 || pitch % 4 || pitch < w || n < 0)
        return fail(TVL1_ERR_INVALID, "bad argument");
    if (n > TVL1_MAX_LEVELS * TVL1_MAX_WARPS) return fail(TVL1_ERR_INVALID, "n too large");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nb = iterate_max_blocks();
    char* tmp = nullptr;
    const size_t part_off = 1024 * ((sizeof(Ctrl) + 1023) / 1024);
    const size_t log_off = part_off + nb * sizeof(double);
    const size_t total = log_off + (size_t)(n + 1) * sizeof(double);
    CK(cudaMalloc(&tmp, total));
    cudaError_t e = cudaMemsetAsync(tmp, 0, total, st);
    if (e != cudaSuccess) { cudaFree(tmp); return fail(TVL1_ERR_CUDA, "memset: %s", cudaGetErrorString(e)); }
    GammaArgs a;
    a.I1wx = d_I1wx; a.I1wy = d_I1wy; a.rho_c = d_rho_c;
    for (int k = 0; k < 2; k++) {
        a.u1[k] = d_u1; a.u2[k] = d_u2; a.p11[k] = d_p11; a.p12[k] = d_p12; a.p21[k] = d_p21; a.p22[k] = d_p22;
    }
    a.u3 = d_u3; a.p31 = d_p31; a.p32 = d_p32;
    a.w = w; a.h = h; a.pitch = pitch; a.l_t = l_t; a.theta = theta; a.taut = taut; a.gamma = gamma;
    a.scaled_eps = -1.f; a.level = 0; a.slot = 0; a.inner_max = 1 << 30; a.mode = 0;
    a.ctrl = (Ctrl*)tmp; a.partials = (double*)(tmp + part_off); a.errlog = (double*)(tmp + log_off);
    int rc = TVL1_OK;
    stage_begin(st);
    for (int i = 0; i < n && !rc; i++) rc = launch_gamma_iteration(a, st);
    stage_end(st);
    if (!rc && errors && n > 0 &&
        (e = cudaMemcpyAsync(errors, tmp + log_off, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess)
        rc = fail(TVL1_ERR_CUDA, "error log copy failed: %s", cudaGetErrorString(e));
    e = cudaStreamSynchronize(st);
    if (!rc && e != cudaSuccess) rc = fail(TVL1_ERR_CUDA, "iterate (gamma): %s", cudaGetErrorString(e));
    cudaFree(tmp);
    return rc;
}

int tvl1_k_median5(const float* d_src, int w, int h, int pitch, float* d_dst, void* stream)
{
    if (!d_src || !d_dst || d_src == d_dst || w <= 0 || h <= 0 || pitch % 4 || pitch < w)
        return fail(TVL1_ERR_INVALID, "bad argument");
    MedianArgs a;
    a.u1[0] = const_cast<float*>(d_src); a.u1[1] = d_dst; a.u2[0] = a.u2[1] = nullptr;
    a.w = w; a.h = h; a.pitch = pitch; a.level = -1; a.slot = 0; a.ctrl = nullptr;
    memset(a.tm, 0, sizeof(a.tm));
    if (int r = make_plane_map(&a.tm[0][0], d_src, pitch, h, TVL1_MED_SW, TVL1_MED_SH)) return r;
    stage_begin((cudaStream_t)stream);
    const int rc = launch_median(a, 1, (cudaStream_t)stream);
    stage_end((cudaStream_t)stream);
    return rc;
}

int tvl1_k_median3(const float* d_src, int w, int h, int pitch, float* d_dst, void* stream)
{
    if (!d_src || !d_dst || d_src == d_dst || w <= 0 || h <= 0 || pitch % 4 || pitch < w)
        return fail(TVL1_ERR_INVALID, "bad argument");
    MedianArgs a;
    a.u1[0] = const_cast<float*>(d_src); a.u1[1] = d_dst; a.u2[0] = a.u2[1] = nullptr;
    a.w = w; a.h = h; a.pitch = pitch; a.level = -1; a.slot = 0; a.ctrl = nullptr;
    memset(a.tm, 0, sizeof(a.tm));
    stage_begin((cudaStream_t)stream);
    const int rc = launch_median3(a, 1, (cudaStream_t)stream);
    stage_end((cudaStream_t)stream);
    return rc;
}

int tvl1_selftest_arith(long long n, unsigned seed, int elo, int ehi, long long* mismatches, long long* unvouched)
{
    if (!mismatches || n <= 0 || elo > ehi || elo < -126 || ehi > 127) return fail(TVL1_ERR_INVALID, "bad argument");
    unsigned long long* d = nullptr;
    CK(cudaMalloc(&d, 2 * sizeof(*d)));
    cudaMemset(d, 0, 2 * sizeof(*d));
    k_selftest_arith<<<148 * 8, 256>>>(seed, n, elo, ehi, d);
    unsigned long long hres[2] = {0, 0};
    cudaError_t e = cudaMemcpy(hres, d, sizeof(hres), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(TVL1_ERR_CUDA, "selftest: %s", cudaGetErrorString(e));
    *mismatches = (long long)hres[0];
    if (unvouched) *unvouched = (long long)hres[1];
    return TVL1_OK;
}

int tvl1_pyramid_sizes(int w, int h, int nscales, double scale_step, int* ws, int* hs)
{
    if (!ws || !hs || w <= 0 || h <= 0 || nscales <= 0) return fail(TVL1_ERR_INVALID, "bad argument");
    return pyramid_sizes(w, h, nscales, scale_step, ws, hs);
}

// ---- 8-bit prescale of the loader (reference src/optflow.cpp:111,124)

static int scaled_size_i(int n, double f) { return (int)lrint((double)n * f); }   // cvRound

int tvl1_prescaled_size(int w, int h, double scale, int* dw, int* dh)
{
    if (!dw || !dh || w <= 0 || h <= 0 || !(scale > 0.0)) return fail(TVL1_ERR_INVALID, "bad argument");
    *dw = scaled_size_i(w, scale);
    *dh = scaled_size_i(h, scale);
    if (*dw <= 0 || *dh <= 0) return fail(TVL1_ERR_INVALID, "scale %g leaves no pixels of %dx%d", scale, w, h);
    return TVL1_OK;
}

int tvl1_prescale_u8(const uint8_t* d_src, size_t spitch, int w, int h, double scale, uint8_t* d_dst, size_t dpitch,
                     void* stream)
{
    int dw = 0, dh = 0;
    int rc = tvl1_prescaled_size(w, h, scale, &dw, &dh);
    if (rc) return rc;
    if (!d_src || !d_dst || spitch < (size_t)w || dpitch < (size_t)dw) return fail(TVL1_ERR_INVALID, "bad buffer or pitch");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 b(32, 8);
    const double inv = 1.0 / scale;
    if (inv == 2.0) k_prescale_half_u8<<<grid2d(dw, dh, b), b, 0, st>>>(d_src, spitch, w, h, d_dst, dpitch, dw, dh);
    else k_prescale_u8<<<grid2d(dw, dh, b), b, 0, st>>>(d_src, spitch, w, h, inv, d_dst, dpitch, dw, dh);
    CK(cudaGetLastError());
    return TVL1_OK;
}

int tvl1_prescale_u8_host(int device, const uint8_t* src, size_t spitch, int w, int h, double scale, uint8_t* dst,
                          size_t dpitch)
{
    int dw = 0, dh = 0;
    int rc = tvl1_prescaled_size(w, h, scale, &dw, &dh);
    if (rc) return rc;
    if (!src || !dst || spitch < (size_t)w || dpitch < (size_t)dw) return fail(TVL1_ERR_INVALID, "bad buffer or pitch");
    if (device < 0 || device >= 64) return fail(TVL1_ERR_INVALID, "device out of range");
    CK(cudaSetDevice(device));
    // grow-only scratch per device (this entry point has no handle to keep it in): no cudaMalloc per call
    struct Scratch { uint8_t* p = nullptr; size_t cap = 0; };
    static Scratch scratch[64];
    static std::mutex mtx;
    std::lock_guard<std::mutex> lock(mtx);
    Scratch& S = scratch[device];
    const size_t raw = ((size_t)w * h + 255) / 256 * 256, need = raw + (size_t)dw * dh;
    if (S.cap < need) {
        if (S.p) cudaFree(S.p);
        S.p = nullptr; S.cap = 0;
        CK(cudaMalloc(&S.p, need));
        S.cap = need;
    }
    uint8_t *ds = S.p, *dd = S.p + raw;
    if (cudaMemcpy2D(ds, w, src, spitch, w, h, cudaMemcpyHostToDevice) != cudaSuccess) return fail(TVL1_ERR_CUDA, "upload failed");
    if ((rc = tvl1_prescale_u8(ds, w, w, h, scale, dd, dw, nullptr))) return rc;
    if (cudaMemcpy2D(dst, dpitch, dd, dw, dw, dh, cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(TVL1_ERR_CUDA, "download failed: %s", cudaGetErrorString(cudaGetLastError()));
    return TVL1_OK;
}

// ---- device-memory helpers

int tvl1_dev_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int tvl1_dev_alloc(int device, size_t bytes, void** out)
{
    if (!out) return fail(TVL1_ERR_INVALID, "out is null");
    CK(cudaSetDevice(device));
    CK(cudaMalloc(out, bytes ? bytes : 1));
    return TVL1_OK;
}

int tvl1_dev_free(int device, void* p)
{
    CK(cudaSetDevice(device));
    CK(cudaFree(p));
    return TVL1_OK;
}

int tvl1_dev_memset(void* d, int v, size_t bytes) { CK(cudaMemset(d, v, bytes)); return TVL1_OK; }
int tvl1_dev_h2d(void* d, const void* h, size_t bytes) { CK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice)); return TVL1_OK; }
int tvl1_dev_d2h(void* h, const void* d, size_t bytes) { CK(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost)); return TVL1_OK; }
int tvl1_dev_sync(int device) { CK(cudaSetDevice(device)); CK(cudaDeviceSynchronize()); return TVL1_OK; }
int tvl1_set_device(int device) { CK(cudaSetDevice(device)); return TVL1_OK; }
int tvl1_stream_create(int device, void** out)
{
    if (!out) return fail(TVL1_ERR_INVALID, "out is null");
    CK(cudaSetDevice(device));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *out = (void*)s;
    return TVL1_OK;
}
int tvl1_stream_destroy(void* stream) { CK(cudaStreamDestroy((cudaStream_t)stream)); return TVL1_OK; }
int tvl1_stream_sync(void* stream) { CK(cudaStreamSynchronize((cudaStream_t)stream)); return TVL1_OK; }
// everything enqueued on `signaller` so far must finish before work enqueued on `waiter` after this call
int tvl1_stream_wait(void* waiter, void* signaller)
{
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaError_t r = cudaEventRecord(e, (cudaStream_t)signaller);
    if (r == cudaSuccess) r = cudaStreamWaitEvent((cudaStream_t)waiter, e, 0);
    cudaEventDestroy(e);   // released once the recorded work has completed
    if (r != cudaSuccess) return fail(TVL1_ERR_CUDA, "stream wait: %s", cudaGetErrorString(r));
    return TVL1_OK;
}
int tvl1_stream_query(void* stream)
{
    const cudaError_t r = cudaStreamQuery((cudaStream_t)stream);
    if (r == cudaSuccess) return 1;
    if (r == cudaErrorNotReady) { cudaGetLastError(); return 0; }
    return fail(TVL1_ERR_CUDA, "stream query: %s", cudaGetErrorString(r));
}
int tvl1_dev_h2d_async(void* d, const void* h, size_t bytes, void* stream)
{
    CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return TVL1_OK;
}
int tvl1_dev_d2h_async(void* h, const void* d, size_t bytes, void* stream)
{
    CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return TVL1_OK;
}
int tvl1_host_alloc_pinned(size_t bytes, void** out)
{
    if (!out) return fail(TVL1_ERR_INVALID, "out is null");
    CK(cudaMallocHost(out, bytes ? bytes : 1));
    return TVL1_OK;
}
int tvl1_host_free_pinned(void* p) { CK(cudaFreeHost(p)); return TVL1_OK; }

}  // extern "C"
