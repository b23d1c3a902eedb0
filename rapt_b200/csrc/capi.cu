// capi.cu -- the extern "C" boundary of librapt_b200.so (include/rapt_b200.h).
// Host-side plumbing only: argument checks, device buffers, H2D/D2H copies, work ordering, launches.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <string>
#include <vector>
#include <atomic>
#include <mutex>
#include "../../include/rapt_b200.h"
#include "rapt_launch.h"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <dlfcn.h>
#include <nvrtc.h>
#include "embedded_headers.inc"

static_assert(sizeof(rapt_field_t) == sizeof(rapt::FieldP), "FieldP layout");
static_assert(sizeof(rapt_params_t) == sizeof(rapt::ParamsP), "ParamsP layout");

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
int g_device = -1, g_sms = 0;
int *g_queue = nullptr;          // device work counters (ring of 64 ints so async calls do not collide)
std::atomic<int> g_queue_slot{0};

// what the last rapt_b200_adaptive_advance of this thread did (rapt_b200_adaptive_last_stats)
struct AdaptiveStats {
    long long epochs = 0, launches_p = 0, launches_g = 0, tracer_launches_p = 0, tracer_launches_g = 0;
    long long steps_p = 0, accepted_p = 0, calls_p = 0, steps_g = 0, calls_g = 0;
    double ms_particle = 0, ms_gc = 0, ms_switch = 0, ms_epochs = 0, ms_total = 0;
    std::vector<float> per_epoch;      // (tracers in particle mode, in guiding-centre mode, particle-kernel ms, gc-kernel ms) x epochs
};
thread_local AdaptiveStats g_adaptive_stats;

// scratch for the longest-first work ordering (grown on demand, reused across calls)
struct SortScratch {
    double *key_in = nullptr, *key_out = nullptr;
    int *idx_in = nullptr, *idx_out = nullptr;
    void *tmp = nullptr;
    size_t n = 0, tmp_bytes = 0;
} g_sort;

#define FLAVOUR(strict, fn, ...) ((strict) ? rapt_strict::fn(__VA_ARGS__) : rapt_fast::fn(__VA_ARGS__))

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? RAPT_E_NODEVICE : RAPT_E_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int ensure_init()
{
    if (g_device >= 0) return RAPT_OK;
    return rapt_b200_init(0);
}

// device-pointer entry points: the state must live on the device the library is bound to (its work counters, sort
// scratch and launches are there); anything else is reported instead of faulting inside the kernel
int check_on_bound_device(const void *p, const char *what)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return RAPT_OK; }
    if (at.type == cudaMemoryTypeDevice && at.device != g_device)
        return fail(RAPT_E_ARG, "%s: the state lives on device %d but librapt_b200 is bound to device %d -- call "
                                "rapt_b200_init(%d) (one process per GPU)", what, at.device, g_device, at.device);
    if (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeUnregistered)
        return fail(RAPT_E_ARG, "%s: host pointer passed to a device-pointer entry point", what);
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != g_device) cudaSetDevice(g_device);
    return RAPT_OK;
}

int *next_queue(cudaStream_t s)
{
    int *q = g_queue + (g_queue_slot.fetch_add(1) & 63);
    cudaMemsetAsync(q, 0, sizeof(int), s);
    return q;
}

// RAII device buffer for the host-pointer entry points
// Stream-ordered allocations from the device's default memory pool (release threshold raised in
// rapt_b200_init, so repeated calls reuse the same memory instead of paying cudaMalloc/cudaFree).
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFreeAsync(p, 0); }
    cudaError_t alloc(size_t b) { bytes = b; return b ? cudaMallocAsync(&p, b, 0) : cudaSuccess; }
    template <class T> T *as() { return static_cast<T *>(p); }
};

cudaError_t up(DevBuf &d, const void *h, size_t bytes, cudaStream_t s)
{
    cudaError_t e = d.alloc(bytes);
    if (e != cudaSuccess || !bytes) return e;
    return cudaMemcpyAsync(d.p, h, bytes, cudaMemcpyHostToDevice, s);
}
cudaError_t down(void *h, DevBuf &d, size_t bytes, cudaStream_t s)
{
    if (!h || !bytes) return cudaSuccess;
    return cudaMemcpyAsync(h, d.p, bytes, cudaMemcpyDeviceToHost, s);
}

// fields.Grid tables resident on the device (rapt_b200_grid_create)
struct GridEntry { rapt::GridP g; bool live = false; void *mem[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; };   // mem[6]: device copy of g
std::vector<GridEntry> g_grids;
std::mutex g_grid_mu;

// rapt_field_t -> FieldP; a gridded field gets its table block placed in prm
int resolve_field(const rapt_field_t *f, rapt::FieldP *out)
{
    memcpy(out, f, sizeof *out);
    if (f->kind == RAPT_FIELD_GRID) {
        std::lock_guard<std::mutex> lk(g_grid_mu);
        if (f->user_id < 0 || f->user_id >= (int)g_grids.size() || !g_grids[f->user_id].live)
            return fail(RAPT_E_ARG, "rapt_field_t.user_id %d is not a live grid handle (rapt_b200_grid_create)", f->user_id);
        memset(out->prm, 0, sizeof out->prm);
        const void *dev = g_grids[f->user_id].mem[6];
        memcpy(&out->prm[0], &dev, sizeof dev);                       // prm[0]: device address of the GridP block
        out->prm[1] = g_grids[f->user_id].g.E ? 1.0 : 0.0;            // prm[1]: has an electric-field table
        out->prm[2] = (double)g_grids[f->user_id].g.nt;               // prm[2]: number of time points (launchers size the cell cache)
    }
    return RAPT_OK;
}
#define RESOLVE(dst, f) do { if (int rc_ = resolve_field((f), &(dst))) return rc_; } while (0)

int fill_common(rapt::AdvArgs &a, const rapt_field_t *f, const rapt_params_t *p)
{
    memset(&a, 0, sizeof a);
    memcpy(&a.p, p, sizeof a.p);
    return resolve_field(f, &a.f);
}


// ------------------------------------------------------------------------------------------------
// NVRTC path for user-defined analytic fields (rapt/fields.py plugin interface; examples/Creating new
// fields.ipynb).  The user's snippet (rapt_user_B / rapt_user_E) is compiled together with the SAME
// kernel templates as the built-ins (headers embedded in this library), for sm_100a, once per
// (source, arithmetic flavour); the cubin is loaded with cudaLibraryLoadData.  libnvrtc is dlopen'ed on
// first use so that the library itself loads on machines without the CUDA toolkit's NVRTC.
// ------------------------------------------------------------------------------------------------
enum { UK_PARTICLE = 0, UK_GC, UK_PARTICLE_DT, UK_FIELD_OPS, UK_MISC, UK_BOUNCE, UK_ADAPT, UK_BC, UK_COUNT };
const char *const k_user_kernel_exprs[UK_COUNT] = {
    "rapt_user::k_particle_dop853<rapt_user::Field<100> >",
    "rapt_user::k_gc_dopri5<rapt_user::Field<100>, 2>",
    "rapt_user::k_particle_dt<rapt_user::Field<100> >",
    "rapt_user::k_field_ops<rapt_user::Field<100> >",
    "rapt_user::k_misc<rapt_user::Field<100> >",
    "rapt_user::k_bounce_setup<rapt_user::Field<100> >",
    "rapt_user::k_adaptive_switch<rapt_user::Field<100> >",
    "rapt_user::k_bounce_center<rapt_user::Field<100> >",
};
struct UserModule { bool built = false; cudaLibrary_t lib = nullptr; cudaKernel_t k[UK_COUNT] = {}; };
struct UserField { std::string src; int has_E = 0; UserModule mod[2]; };
std::vector<UserField> g_user;

struct Nvrtc {
    void *h = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram *) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char *) = nullptr;
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char *, const char **) = nullptr;
    const char *(*GetErrorString)(nvrtcResult) = nullptr;
} g_nvrtc;

int load_nvrtc()
{
    if (g_nvrtc.h) return RAPT_OK;
    const char *cands[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/targets/x86_64-linux/lib/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char *c : cands) { g_nvrtc.h = dlopen(c, RTLD_NOW | RTLD_GLOBAL); if (g_nvrtc.h) break; }
    if (!g_nvrtc.h) return fail(RAPT_E_NVRTC, "cannot dlopen libnvrtc (%s)", dlerror());
#define SYM(n) *(void **)(&g_nvrtc.n) = dlsym(g_nvrtc.h, "nvrtc" #n); if (!g_nvrtc.n) return fail(RAPT_E_NVRTC, "libnvrtc lacks nvrtc" #n)
    SYM(CreateProgram); SYM(DestroyProgram); SYM(CompileProgram); SYM(GetProgramLogSize); SYM(GetProgramLog);
    SYM(GetCUBINSize); SYM(GetCUBIN); SYM(AddNameExpression); SYM(GetLoweredName); SYM(GetErrorString);
#undef SYM
    return RAPT_OK;
}

// compile (source, flavour) to a cubin; returns the lowered kernel names
int nvrtc_compile(const UserField &uf, bool strict, std::string &cubin, std::string lowered[UK_COUNT], char *log, int loglen)
{
    if (int rc = load_nvrtc()) return rc;
    std::string src;
    src += "#define RAPT_USER_FIELD 1\n";
    src += std::string("#define RAPT_USER_HAS_E ") + (uf.has_E ? "1" : "0") + "\n";
    src += std::string("#define RAPT_STRICT ") + (strict ? "1" : "0") + "\n";
    src += "#define RAPT_NS rapt_user\n";
    src += "#define RAPT_GC_THREADS 128\n";           // user kernels are launched with 128-thread blocks (launch_any)
    src += "#include \"rapt_bc.cuh\"\n";
    src += "namespace rapt_user {\n"
           "template <class F> __global__ void k_particle_dt(const rapt::AdvArgs a, double *key, int *idx) {\n"
           "  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; if (i >= a.nwork) return;\n"
           "  double mass = a.mass[i], q = a.charge[i], px = a.s4[i], py = a.s5[i], pz = a.s6[i];\n"
           "  double gm = sqrt(mass * mass + dot3(px, py, pz, px, py, pz) / (RAPT_C_LIGHT * RAPT_C_LIGHT));\n"
           "  double vx = px / gm, vy = py / gm, vz = pz / gm;\n"
           "  double gamma = 1.0 / sqrt(1 - dot3(vx, vy, vz, vx, vy, vz) / (RAPT_C_LIGHT * RAPT_C_LIGHT));\n"
           "  double Bm = F::magB(a.f, a.t[i], a.s1[i], a.s2[i], a.s3[i]);\n"
           "  double dt = 2 * RAPT_PI * gamma * mass / Bm / fabs(q) / a.p.cyclotronresolution;\n"
           "  double delta = a.delta_arr ? a.delta_arr[i] : a.delta;\n"
           "  key[i] = dt / delta; idx[i] = (int)i; }\n}\n";
    src += "// ---- user snippet\nnamespace rapt_user {\n" + uf.src + "\n}\n";
    nvrtcProgram prog;
    nvrtcResult r = g_nvrtc.CreateProgram(&prog, src.c_str(), "rapt_user_field.cu", k_nhdr, k_hdr_srcs, k_hdr_names);
    if (r != NVRTC_SUCCESS) return fail(RAPT_E_NVRTC, "nvrtcCreateProgram: %s", g_nvrtc.GetErrorString(r));
    for (int k = 0; k < UK_COUNT; k++) g_nvrtc.AddNameExpression(prog, k_user_kernel_exprs[k]);
    std::vector<const char *> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
    if (strict) opts.push_back("--fmad=false");
    r = g_nvrtc.CompileProgram(prog, (int)opts.size(), opts.data());
    size_t ls = 0;
    g_nvrtc.GetProgramLogSize(prog, &ls);
    if (ls > 1 && log && loglen > 0) {
        std::string l(ls, '\0');
        g_nvrtc.GetProgramLog(prog, &l[0]);
        snprintf(log, loglen, "%s", l.c_str());
    }
    if (r != NVRTC_SUCCESS) {
        g_nvrtc.DestroyProgram(&prog);
        return fail(RAPT_E_NVRTC, "nvrtcCompileProgram: %s", g_nvrtc.GetErrorString(r));
    }
    for (int k = 0; k < UK_COUNT; k++) {
        const char *ln = nullptr;
        if (g_nvrtc.GetLoweredName(prog, k_user_kernel_exprs[k], &ln) != NVRTC_SUCCESS || !ln) {
            g_nvrtc.DestroyProgram(&prog);
            return fail(RAPT_E_NVRTC, "no lowered name for %s", k_user_kernel_exprs[k]);
        }
        lowered[k] = ln;
    }
    size_t cs = 0;
    g_nvrtc.GetCUBINSize(prog, &cs);
    cubin.resize(cs);
    g_nvrtc.GetCUBIN(prog, &cubin[0]);
    g_nvrtc.DestroyProgram(&prog);
    return RAPT_OK;
}

int user_module(int uid, bool strict, UserModule **out)
{
    if (uid < 0 || uid >= (int)g_user.size()) return fail(RAPT_E_ARG, "unknown user field id %d", uid);
    UserModule &m = g_user[uid].mod[strict ? 1 : 0];
    if (!m.built) {
        std::string cubin, lowered[UK_COUNT];
        char log[4096] = "";
        if (int rc = nvrtc_compile(g_user[uid], strict, cubin, lowered, log, sizeof log)) return rc;
        CK(cudaLibraryLoadData(&m.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
        for (int k = 0; k < UK_COUNT; k++) CK(cudaLibraryGetKernel(&m.k[k], m.lib, lowered[k].c_str()));
        m.built = true;
    }
    *out = &m;
    return RAPT_OK;
}

// One launch of kernel family `which` for field f: built-in -> the precompiled flavour; user -> NVRTC module.
int launch_any(const rapt_field_t *f, bool strict, int which, const void *args, long long n, int grid, cudaStream_t s,
               double *key = nullptr, int *idx = nullptr)
{
    if (f->kind == RAPT_FIELD_USER) {
        UserModule *m = nullptr;
        if (int rc = user_module(f->user_id, strict, &m)) return rc;
        int block = 128;
        if (which == UK_PARTICLE_DT) { block = 256; grid = (int)((n + 255) / 256); }
        else if (which == UK_ADAPT) { block = 256; grid = (int)((n + 255) / 256); }
        else if (which == UK_FIELD_OPS || which == UK_MISC || which == UK_BOUNCE) grid = (int)((n + 127) / 128);
        else if (which == UK_BC) block = 64;
        void *kargs[3] = {const_cast<void *>(args), &key, &idx};
        CK(cudaLaunchKernel((const void *)m->k[which], dim3(grid), dim3(block), kargs, 0, s));
        return RAPT_OK;
    }
    switch (which) {
    case UK_PARTICLE: CK(FLAVOUR(strict, launch_particle, *static_cast<const rapt::AdvArgs *>(args), grid, s)); break;
    case UK_GC: CK(FLAVOUR(strict, launch_gc, *static_cast<const rapt::AdvArgs *>(args), grid, s)); break;
    case UK_PARTICLE_DT: CK(FLAVOUR(strict, launch_particle_dt, *static_cast<const rapt::AdvArgs *>(args), key, idx, s)); break;
    case UK_FIELD_OPS: CK(FLAVOUR(strict, launch_field_ops, args, s)); break;
    case UK_MISC: CK(FLAVOUR(strict, launch_misc, args, s)); break;
    case UK_BOUNCE: CK(FLAVOUR(strict, launch_bounce, args, s)); break;
    case UK_ADAPT: CK(FLAVOUR(strict, launch_adaptive_switch, args, s)); break;
    case UK_BC: CK(FLAVOUR(strict, launch_bounce_center, args, grid, s)); break;
    default: return fail(RAPT_E_ARG, "bad kernel family");
    }
    return RAPT_OK;
}

int check_field(const rapt_field_t *f)
{
    if (!f) return fail(RAPT_E_ARG, "null field");
    if (f->kind == RAPT_FIELD_USER) {
        if (f->user_id < 0 || f->user_id >= (int)g_user.size()) return fail(RAPT_E_ARG, "unknown user field id %d", f->user_id);
        return RAPT_OK;
    }
    if (f->kind < 0 || f->kind > RAPT_FIELD_GRID) return fail(RAPT_E_ARG, "unknown field kind %d", f->kind);
    return RAPT_OK;
}

// Work-order key of a guiding centre: 1 / (expected steps of this call).  A-priori estimate: DOPRI5 resolves the parallel
// (bounce) motion, whose time scale is the transit time |r| / v of the tracer across its own distance from the origin of a
// planet-centred field; ~23 steps per transit at the default tolerances on configs 3 and 5 (oracle counts, correlation 0.93 /
// 0.96; list scheduling 2.2 % / 1.3 % above the ideal makespan against 8.5 % in member order: profiles/r2_work_order.md).
// use_prev (rapt_params_t.sort_by_work = 2): the tracer's own step count of the previous call, left in the counters buffer
// by the caller, replaces the estimate (0: no history).  Scheduling only.
__global__ void __launch_bounds__(256) k_key_gc(long long n, const double *__restrict__ x, const double *__restrict__ y,
                                               const double *__restrict__ z, const double *__restrict__ v,
                                               const double *__restrict__ delta_arr, double delta,
                                               const int *__restrict__ counters, int use_prev, double *key, int *idx)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r = sqrt(x[i] * x[i] + y[i] * y[i] + z[i] * z[i]);
    double est = 23.0 * (delta_arr ? delta_arr[i] : delta) * v[i] / r;
    if (!(est > 1.0) || est > 1e15) est = 1.0;                  // NaN, r = 0, v = 0: no estimate
    if (use_prev) { const int prev = counters[4 * i + 1]; if (prev > 0) est = (double)prev; }
    key[i] = 1.0 / est;
    idx[i] = (int)i;
}

// Longest-first schedule: sort particle indices by dt/delta ascending (fewest-rows last), so the lanes
// that run dry at the end of the kernel are finishing the SHORTEST particles (SURVEY.md hard part H4).
// gc: the guiding-centre key (k_key_gc) instead of the Particle one.
int build_order(const rapt_field_t *f, rapt::AdvArgs &a, bool strict, cudaStream_t s, bool gc = false)
{
    const size_t n = (size_t)a.nwork;
    if (g_sort.n < n) {
        cudaFree(g_sort.key_in); cudaFree(g_sort.key_out); cudaFree(g_sort.idx_in); cudaFree(g_sort.idx_out);
        g_sort.key_in = g_sort.key_out = nullptr; g_sort.idx_in = g_sort.idx_out = nullptr; g_sort.n = 0;   // a failed cudaMalloc below must not leave freed pointers behind
        CK(cudaMalloc(&g_sort.key_in, n * sizeof(double)));
        CK(cudaMalloc(&g_sort.key_out, n * sizeof(double)));
        CK(cudaMalloc(&g_sort.idx_in, n * sizeof(int)));
        CK(cudaMalloc(&g_sort.idx_out, n * sizeof(int)));
        g_sort.n = n;
    }
    if (gc) {
        k_key_gc<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((long long)n, a.s1, a.s2, a.s3, a.v, a.delta_arr, a.delta, a.counters,
                                                             a.p.sort_by_work == 2 && !a.append, g_sort.key_in, g_sort.idx_in);
        CK(cudaGetLastError());
    } else if (int rc = launch_any(f, strict, UK_PARTICLE_DT, &a, (long long)n, 0, s, g_sort.key_in, g_sort.idx_in)) return rc;
    g_launches++;
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, g_sort.key_in, g_sort.key_out, g_sort.idx_in, g_sort.idx_out, (int)n, 0, 64, s);
    if (need > g_sort.tmp_bytes) {
        cudaFree(g_sort.tmp);
        g_sort.tmp = nullptr; g_sort.tmp_bytes = 0;
        CK(cudaMalloc(&g_sort.tmp, need));
        g_sort.tmp_bytes = need;
    }
    CK(cub::DeviceRadixSort::SortPairs(g_sort.tmp, need, g_sort.key_in, g_sort.key_out, g_sort.idx_in, g_sort.idx_out, (int)n, 0, 64, s));
    g_launches += 1;
    a.order = g_sort.idx_out;
    return RAPT_OK;
}

int grid_for(long long n, int blocks_per_sm)
{
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    long long want = (n + 127) / 128;
    return (int)std::min<long long>((long long)g_sms * blocks_per_sm, want);
}

// register-resident DFMA chains: 8 independent accumulators per thread, 2 flops per DFMA
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678) out[0] = s;   // never true; keeps the chains alive
}

// Final-state diagnostics of a shard in ONE pass over the state columns (BASELINE.json north_star: "all-gather final
// states and diagnostics (histograms, invariants)"): packs the SoA columns into the [n][ncol] rows the all-gather sends,
// bins a per-tracer quantity into a shared-memory histogram (one global atomic per bin and block) and accumulates the
// sums of an invariant.  kind 0: log10 of the kinetic energy in eV from (px, py, pz, mass) -- Particle.getke,
// Particle.py:442-454; kind 1: radial distance in Earth radii (drift-shell occupation of a guiding-centre ensemble).
// stats: [0] tracers with status 1, [1] sum q, [2] sum q^2, [3] tracers outside [lo, hi).
__global__ void __launch_bounds__(256) k_final_diagnostics(int kind, long long n, int ncol, rapt::DiagCols c,
                                                           const double *__restrict__ mass, const int *__restrict__ status,
                                                           double *__restrict__ packed, int nbins, double lo, double hi,
                                                           unsigned long long *__restrict__ hist, double *__restrict__ stats)
{
    extern __shared__ unsigned int sh_hist[];
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) sh_hist[b] = 0;
    __syncthreads();
    double s_ok = 0, s_q = 0, s_q2 = 0, s_out = 0;
    const double scale = nbins / (hi - lo);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (k < ncol) ? c.col[k][i] : 0.0;
        if (packed) {
            double *row = packed + i * ncol;
#pragma unroll
            for (int k = 0; k < 8; k++) if (k < ncol) row[k] = v[k];
        }
        double q;
        if (kind == 0) {
            const double m = mass[i], mc = m * 299792458.0;
            const double p2 = v[4] * v[4] + v[5] * v[5] + v[6] * v[6];
            const double g = sqrt(1.0 + p2 / (mc * mc));
            const double ke = (g - 1.0 < 1e-6) ? 0.5 * p2 / m : (g - 1.0) * mc * 299792458.0;
            q = log10(ke / 1.602176565e-19);
        } else {
            q = sqrt(v[1] * v[1] + v[2] * v[2] + v[3] * v[3]) / 6378137.0;
        }
        const bool ok = !status || status[i] == 1;
        if (ok) { s_ok += 1; s_q += q; s_q2 += q * q; }
        const int b = (int)floor((q - lo) * scale);
        if (q >= lo && b < nbins) atomicAdd(&sh_hist[b], 1u);
        else if (q == hi) atomicAdd(&sh_hist[nbins - 1], 1u);        // closed last bin, as numpy.histogram
        else s_out += 1;
    }
    // warp-reduce the four sums, one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_ok += __shfl_xor_sync(0xffffffffu, s_ok, o); s_q += __shfl_xor_sync(0xffffffffu, s_q, o);
        s_q2 += __shfl_xor_sync(0xffffffffu, s_q2, o); s_out += __shfl_xor_sync(0xffffffffu, s_out, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&stats[0], s_ok); atomicAdd(&stats[1], s_q); atomicAdd(&stats[2], s_q2); atomicAdd(&stats[3], s_out);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x)
        if (sh_hist[b]) atomicAdd(&hist[b], (unsigned long long)sh_hist[b]);
}

// all-gathered rows [world][n_max][ncol] -> member order [n_total][ncol].  Shards are periodic: of every `period` consecutive
// members, those at positions [off[r], off[r+1]) belong to rank r (round-robin is period = world, one position each;
// speed-weighted shards use a longer period with unequal runs: rapt_b200/dist.py:ShardPlan); the last, partial period
// (block index `full`) is cut at off2 (the same proportions scaled to its length).  Coalesced writes.
struct ShardTable { int period, world; long long full; int off[65], off2[65]; };
__global__ void __launch_bounds__(256) k_unshard(ShardTable tb, long long n_max, int ncol, long long n_total,
                                                const double *__restrict__ buf, double *__restrict__ out)
{
    const long long total = n_total * ncol;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long m = e / ncol; const int c = (int)(e - m * ncol);
        const long long blk = m / tb.period; const int j = (int)(m - blk * tb.period);
        const int *o = (blk == tb.full) ? tb.off2 : tb.off;
        int r = 0;
        while (r + 1 < tb.world && j >= o[r + 1]) r++;
        const long long i = blk * (tb.off[r + 1] - tb.off[r]) + (j - o[r]);
        out[e] = buf[((long long)r * n_max + i) * ncol + c];
    }
}

}  // namespace

extern "C" {

const char *rapt_b200_last_error(void) { return g_err.c_str(); }
const char *rapt_b200_version(void) { return "rapt_b200 0.1 (sm_100a)"; }
int64_t rapt_b200_launch_count(void) { return g_launches.load(); }

int rapt_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int rapt_b200_init(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(RAPT_E_NODEVICE, "no CUDA device available (%s); librapt_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(RAPT_E_ARG, "device %d out of range [0,%d)", device, n);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (g_device != device) {
        // (re)bind the library to this device: the work counters and the sort scratch live in its memory.  One process
        // drives one GPU; a later call with another device moves the binding, it does not share it.
        g_queue = nullptr;
        CK(cudaMalloc(&g_queue, 64 * sizeof(int)));
        CK(cudaMemset(g_queue, 0, 64 * sizeof(int)));
        g_sort = SortScratch();
    }
    g_device = device;
    g_sms = prop.multiProcessorCount;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return RAPT_OK;
}


int rapt_b200_fp64_peak(int iters, double *tflops, double *sm_clock_mhz)
{
    if (int rc = ensure_init()) return rc;
    if (iters <= 0 || !tflops) return fail(RAPT_E_ARG, "fp64_peak: bad argument");
    double *d = nullptr;
    CK(cudaMalloc(&d, sizeof(double)));
    const int blocks = g_sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_fp64_peak<<<blocks, threads>>>(d, iters / 8 + 1, 0.999999, 1e-9);        // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0));
        k_fp64_peak<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
        g_launches++;
    }
    CK(cudaGetLastError());
    double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (sm_clock_mhz) {
        int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g_device);
        *sm_clock_mhz = khz / 1000.0;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return RAPT_OK;
}

int rapt_b200_field_nvrtc(const char *cuda_src, int has_E, int *user_id, char *log, int loglen)
{
    if (!cuda_src || !user_id) return fail(RAPT_E_ARG, "field_nvrtc: null argument");
    if (log && loglen > 0) log[0] = 0;
    for (size_t i = 0; i < g_user.size(); i++)
        if (g_user[i].has_E == (has_E != 0) && g_user[i].src == cuda_src) { *user_id = (int)i; return RAPT_OK; }
    UserField uf;
    uf.src = cuda_src; uf.has_E = has_E != 0;
    // compile the fast flavour now so that syntax errors surface here (no device needed for NVRTC itself)
    std::string cubin, lowered[UK_COUNT];
    if (int rc = nvrtc_compile(uf, false, cubin, lowered, log, loglen)) return rc;
    g_user.push_back(uf);
    *user_id = (int)g_user.size() - 1;
    if (rapt_b200_device_count() > 0 && ensure_init() == RAPT_OK) {
        UserModule &m = g_user.back().mod[0];
        CK(cudaLibraryLoadData(&m.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
        for (int k = 0; k < UK_COUNT; k++) CK(cudaLibraryGetKernel(&m.k[k], m.lib, lowered[k].c_str()));
        m.built = true;
    }
    return RAPT_OK;
}

int rapt_b200_grid_create(int64_t nt, int64_t nx, int64_t ny, int64_t nz,
                          const double *t, const double *x, const double *y, const double *z,
                          const double *Bx, const double *By, const double *Bz,
                          const double *Ex, const double *Ey, const double *Ez, int32_t *grid_id)
{
    if (int rc = ensure_init()) return rc;
    if (nt < 1 || nx < 2 || ny < 2 || nz < 2 || !x || !y || !z || !Bx || !By || !Bz || !grid_id || (nt > 1 && !t))
        return fail(RAPT_E_ARG, "rapt_b200_grid_create: need nt >= 1, nx, ny, nz >= 2 and the coordinate and B arrays");
    if ((Ex || Ey || Ez) && !(Ex && Ey && Ez)) return fail(RAPT_E_ARG, "rapt_b200_grid_create: give all of Ex, Ey, Ez or none");
    const double *axes[4] = {t, x, y, z};
    const int64_t len[4] = {nt, nx, ny, nz};
    for (int a = (nt > 1 ? 0 : 1); a < 4; a++)
        for (int64_t i = 1; i < len[a]; i++)
            if (!(axes[a][i] > axes[a][i - 1])) return fail(RAPT_E_ARG, "rapt_b200_grid_create: axis %d is not strictly ascending", a);
    GridEntry e;
    rapt::GridP &g = e.g;
    memset(&g, 0, sizeof g);
    g.nt = (int)nt; g.nx = (int)nx; g.ny = (int)ny; g.nz = (int)nz;
    // coordinates
    const double **dst[4] = {&g.t, &g.x, &g.y, &g.z};
    for (int a = 0; a < 4; a++) {
        if (a == 0 && nt == 1 && !t) continue;
        CK(cudaMalloc(&e.mem[a], len[a] * sizeof(double)));
        CK(cudaMemcpy(e.mem[a], axes[a], len[a] * sizeof(double), cudaMemcpyHostToDevice));
        *dst[a] = static_cast<const double *>(e.mem[a]);
    }
    // uniform axes get a direct index (made exact against the node array on the device)
    double *u0[3] = {&g.x0, &g.y0, &g.z0}, *ui[3] = {&g.xinv, &g.yinv, &g.zinv};
    for (int a = 1; a < 4; a++) {
        const double *v = axes[a]; const int64_t n = len[a];
        const double d = (v[n - 1] - v[0]) / (double)(n - 1);
        bool uniform = true;
        for (int64_t i = 0; i < n && uniform; i++) uniform = fabs(v[i] - (v[0] + d * (double)i)) <= 1e-9 * fabs(d);
        *u0[a - 1] = v[0]; *ui[a - 1] = uniform ? 1.0 / d : 0.0;
    }
    // node tables: [nt][nx][ny][nz] x (c0, c1, c2, pad), staged one time point at a time
    const size_t nodes = (size_t)nx * ny * nz;
    bool hasE = false;
    if (Ex) for (size_t i = 0; i < nodes * (size_t)nt && !hasE; i++) hasE = (Ex[i] != 0.0) || (Ey[i] != 0.0) || (Ez[i] != 0.0);
    std::vector<double> stage(nodes * 4);
    const double *comp[2][3] = {{Bx, By, Bz}, {Ex, Ey, Ez}};
    for (int w = 0; w < 2; w++) {
        if (w == 1 && !hasE) break;
        CK(cudaMalloc(&e.mem[4 + w], nodes * (size_t)nt * 4 * sizeof(double)));
        for (int64_t it = 0; it < nt; it++) {
            const size_t off = (size_t)it * nodes;
            for (size_t i = 0; i < nodes; i++) {
                stage[4 * i] = comp[w][0][off + i]; stage[4 * i + 1] = comp[w][1][off + i];
                stage[4 * i + 2] = comp[w][2][off + i]; stage[4 * i + 3] = 0.0;
            }
            CK(cudaMemcpy(static_cast<double *>(e.mem[4 + w]) + off * 4, stage.data(), nodes * 4 * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    g.B = static_cast<const double *>(e.mem[4]);
    g.E = static_cast<const double *>(e.mem[5]);
    CK(cudaMalloc(&e.mem[6], sizeof g));
    CK(cudaMemcpy(e.mem[6], &g, sizeof g, cudaMemcpyHostToDevice));
    e.live = true;
    std::lock_guard<std::mutex> lk(g_grid_mu);
    g_grids.push_back(e);
    *grid_id = (int32_t)g_grids.size() - 1;
    return RAPT_OK;
}

int rapt_b200_grid_destroy(int32_t grid_id)
{
    std::lock_guard<std::mutex> lk(g_grid_mu);
    if (grid_id < 0 || grid_id >= (int)g_grids.size() || !g_grids[grid_id].live)
        return fail(RAPT_E_ARG, "rapt_b200_grid_destroy: %d is not a live grid handle", grid_id);
    for (void *&m : g_grids[grid_id].mem) { if (m) cudaFree(m); m = nullptr; }
    g_grids[grid_id].live = false;
    return RAPT_OK;
}

// ------------------------------------------------------------------------------------------------
// Particle.advance
// ------------------------------------------------------------------------------------------------
int rapt_b200_particle_advance_dev(const rapt_field_t *f, const rapt_params_t *p, int64_t n,
                                   double *t, double *x, double *y, double *z, double *px, double *py, double *pz,
                                   const double *mass, const double *charge, double delta,
                                   int64_t store_every, int64_t max_rows, double *rows,
                                   int32_t *nrows, int32_t *nstored, int32_t *counters, int32_t *status,
                                   double *tcur, double *dt_out, void *stream)
{
    if (int rc = ensure_init()) return rc;
    if (!f || !p || n < 0 || !t || !x || !y || !z || !px || !py || !pz || !mass || !charge || !nstored || !counters ||
        !status || !tcur)
        return fail(RAPT_E_ARG, "particle_advance: null argument");
    if (int rc = check_field(f)) return rc;
    if (!(p->rtol > 0) || !(p->atol > 0) || !(p->cyclotronresolution > 0))
        return fail(RAPT_E_ARG, "particle_advance: solvertolerances and cyclotronresolution must be positive");
    if (n == 0) return RAPT_OK;
    if (int rc = check_on_bound_device(t, "particle_advance_dev")) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rapt::AdvArgs a;
    if (int rc = fill_common(a, f, p)) return rc;
    a.nwork = n; a.order = nullptr; a.queue = next_queue(s);
    a.t = t; a.s1 = x; a.s2 = y; a.s3 = z; a.s4 = px; a.s5 = py; a.s6 = pz;
    a.mass = mass; a.charge = charge; a.delta = delta;
    a.store_every = rows ? store_every : 0; a.max_rows = rows ? max_rows : 0; a.rows = rows;
    a.nstored = nstored; a.nrows = nrows; a.counters = counters; a.status = status; a.tcur = tcur; a.dt_out = dt_out;
    const bool strict = p->arith == 1;
    const int rkn = !strict && f->is_static && !p->enforce_equatorial && f->kind != RAPT_FIELD_USER && !getenv("RAPT_B200_NO_RKN");
    const int grid = grid_for(n, FLAVOUR(strict, particle_blocks_per_sm, rkn));
    if (p->sort_by_work && n > (long long)grid * 128) {
        if (int rc = build_order(f, a, strict, s)) return rc;
        if (rkn && getenv("RAPT_B200_SPREAD")) a.spread_first_wave = grid * 128;        // opt-in: measured, no gain (profiles/r2_tail.md)
    }
    if (int rc = launch_any(f, strict, UK_PARTICLE, &a, n, grid, s)) return rc;
    g_launches++;
    return RAPT_OK;
}

int rapt_b200_particle_advance(const rapt_field_t *f, const rapt_params_t *p, int64_t n,
                               double *t, double *x, double *y, double *z, double *px, double *py, double *pz,
                               const double *mass, const double *charge, double delta,
                               int64_t store_every, int64_t max_rows, double *rows,
                               int32_t *nrows, int32_t *nstored, int32_t *counters, int32_t *status,
                               double *tcur, double *dt_out)
{
    if (int rc = ensure_init()) return rc;
    if (n < 0) return fail(RAPT_E_ARG, "n < 0");
    if (n == 0) return RAPT_OK;
    if (!t || !x || !y || !z || !px || !py || !pz || !mass || !charge)
        return fail(RAPT_E_ARG, "particle_advance: null state pointer");
    cudaStream_t s = 0;
    const size_t nb = (size_t)n * sizeof(double);
    DevBuf dt_, dx, dy, dz, dpx, dpy, dpz, dm, dq, drows, dnrows, dnst, dcnt, dst, dtcur, ddt;
    CK(up(dt_, t, nb, s)); CK(up(dx, x, nb, s)); CK(up(dy, y, nb, s)); CK(up(dz, z, nb, s));
    CK(up(dpx, px, nb, s)); CK(up(dpy, py, nb, s)); CK(up(dpz, pz, nb, s));
    CK(up(dm, mass, nb, s)); CK(up(dq, charge, nb, s));
    const bool want_rows = rows && store_every > 0 && max_rows > 0;
    const size_t rb = want_rows ? (size_t)n * (size_t)max_rows * 8 * sizeof(double) : 0;
    CK(drows.alloc(rb));
    CK(dnrows.alloc(n * sizeof(int))); CK(dnst.alloc(n * sizeof(int))); CK(dcnt.alloc(n * 4 * sizeof(int)));
    CK(dst.alloc(n * sizeof(int))); CK(dtcur.alloc(nb)); CK(ddt.alloc(nb));
    rapt_params_t ph = *p;
    if (ph.sort_by_work > 1) ph.sort_by_work = 1;          // no history in freshly allocated output buffers
    int rc = rapt_b200_particle_advance_dev(f, &ph, n, dt_.as<double>(), dx.as<double>(), dy.as<double>(), dz.as<double>(),
                                            dpx.as<double>(), dpy.as<double>(), dpz.as<double>(), dm.as<double>(),
                                            dq.as<double>(), delta, store_every, max_rows,
                                            want_rows ? drows.as<double>() : nullptr, dnrows.as<int>(), dnst.as<int>(),
                                            dcnt.as<int>(), dst.as<int>(), dtcur.as<double>(), ddt.as<double>(), s);
    if (rc) return rc;
    CK(down(t, dt_, nb, s)); CK(down(x, dx, nb, s)); CK(down(y, dy, nb, s)); CK(down(z, dz, nb, s));
    CK(down(px, dpx, nb, s)); CK(down(py, dpy, nb, s)); CK(down(pz, dpz, nb, s));
    CK(down(nrows, dnrows, n * sizeof(int), s)); CK(down(nstored, dnst, n * sizeof(int), s));
    CK(down(counters, dcnt, n * 4 * sizeof(int), s)); CK(down(status, dst, n * sizeof(int), s));
    CK(down(tcur, dtcur, nb, s)); CK(down(dt_out, ddt, nb, s));
    if (want_rows) CK(down(rows, drows, rb, s));
    CK(cudaStreamSynchronize(s));
    return RAPT_OK;
}

// ------------------------------------------------------------------------------------------------
// GuidingCenter.advance
// ------------------------------------------------------------------------------------------------
int rapt_b200_gc_advance_dev(const rapt_field_t *f, const rapt_params_t *p, int eom, int64_t n,
                             double *t, double *x, double *y, double *z, double *ppar,
                             const double *mu, const double *v, const double *mass, const double *charge,
                             const double *dt, double delta,
                             int64_t store_every, int64_t max_rows, double *rows,
                             int32_t *nrows, int32_t *nstored, int32_t *counters, int32_t *status, double *tcur,
                             void *stream)
{
    if (int rc = ensure_init()) return rc;
    if (!f || !p || n < 0 || !t || !x || !y || !z || !ppar || !mu || !v || !mass || !charge || !dt || !nstored ||
        !counters || !status || !tcur)
        return fail(RAPT_E_ARG, "gc_advance: null argument");
    if (int rc = check_field(f)) return rc;
    if (eom < 0 || eom > 2) return fail(RAPT_E_ARG, "unknown eom %d", eom);
    if (!(p->rtol > 0) || !(p->atol > 0)) return fail(RAPT_E_ARG, "gc_advance: solvertolerances must be positive");
    if (n == 0) return RAPT_OK;
    if (int rc = check_on_bound_device(t, "gc_advance_dev")) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    rapt::AdvArgs a;
    if (int rc = fill_common(a, f, p)) return rc;
    a.nwork = n; a.order = nullptr; a.queue = next_queue(s);
    a.t = t; a.s1 = x; a.s2 = y; a.s3 = z; a.s4 = ppar;
    a.mass = mass; a.charge = charge; a.mu = mu; a.v = v; a.dtin = dt; a.delta = delta; a.eom = eom;
    a.store_every = rows ? store_every : 0; a.max_rows = rows ? max_rows : 0; a.rows = rows;
    a.nstored = nstored; a.nrows = nrows; a.counters = counters; a.status = status; a.tcur = tcur;
    const bool strict = p->arith == 1;
    const int grid = grid_for(n, FLAVOUR(strict, gc_blocks_per_sm));
    // longest-first: a-priori estimate (sort_by_work = 1) or the previous call's step counts (2): k_key_gc
    if (p->sort_by_work && n > (long long)grid * 128) {
        if (int rc = build_order(f, a, strict, s, true)) return rc;
    }
    if (int rc = launch_any(f, strict, UK_GC, &a, n, grid, s)) return rc;
    g_launches++;
    return RAPT_OK;
}

int rapt_b200_gc_advance(const rapt_field_t *f, const rapt_params_t *p, int eom, int64_t n,
                         double *t, double *x, double *y, double *z, double *ppar,
                         const double *mu, const double *v, const double *mass, const double *charge,
                         const double *dt, double delta,
                         int64_t store_every, int64_t max_rows, double *rows,
                         int32_t *nrows, int32_t *nstored, int32_t *counters, int32_t *status, double *tcur)
{
    if (int rc = ensure_init()) return rc;
    if (n < 0) return fail(RAPT_E_ARG, "n < 0");
    if (n == 0) return RAPT_OK;
    if (!t || !x || !y || !z || !ppar || !mu || !v || !mass || !charge || !dt)
        return fail(RAPT_E_ARG, "gc_advance: null state pointer");
    cudaStream_t s = 0;
    const size_t nb = (size_t)n * sizeof(double);
    DevBuf dt_, dx, dy, dz, dpp, dmu, dv, dm, dq, ddt, drows, dnrows, dnst, dcnt, dst, dtcur;
    CK(up(dt_, t, nb, s)); CK(up(dx, x, nb, s)); CK(up(dy, y, nb, s)); CK(up(dz, z, nb, s)); CK(up(dpp, ppar, nb, s));
    CK(up(dmu, mu, nb, s)); CK(up(dv, v, nb, s)); CK(up(dm, mass, nb, s)); CK(up(dq, charge, nb, s)); CK(up(ddt, dt, nb, s));
    const bool want_rows = rows && store_every > 0 && max_rows > 0;
    const size_t rb = want_rows ? (size_t)n * (size_t)max_rows * 8 * sizeof(double) : 0;
    CK(drows.alloc(rb));
    CK(dnrows.alloc(n * sizeof(int))); CK(dnst.alloc(n * sizeof(int))); CK(dcnt.alloc(n * 4 * sizeof(int)));
    CK(dst.alloc(n * sizeof(int))); CK(dtcur.alloc(nb));
    rapt_params_t ph = *p;
    if (ph.sort_by_work > 1) ph.sort_by_work = 1;          // no history in freshly allocated output buffers
    int rc = rapt_b200_gc_advance_dev(f, &ph, eom, n, dt_.as<double>(), dx.as<double>(), dy.as<double>(), dz.as<double>(),
                                      dpp.as<double>(), dmu.as<double>(), dv.as<double>(), dm.as<double>(), dq.as<double>(),
                                      ddt.as<double>(), delta, store_every, max_rows,
                                      want_rows ? drows.as<double>() : nullptr, dnrows.as<int>(), dnst.as<int>(),
                                      dcnt.as<int>(), dst.as<int>(), dtcur.as<double>(), s);
    if (rc) return rc;
    CK(down(t, dt_, nb, s)); CK(down(x, dx, nb, s)); CK(down(y, dy, nb, s)); CK(down(z, dz, nb, s)); CK(down(ppar, dpp, nb, s));
    CK(down(nrows, dnrows, n * sizeof(int), s)); CK(down(nstored, dnst, n * sizeof(int), s));
    CK(down(counters, dcnt, n * 4 * sizeof(int), s)); CK(down(status, dst, n * sizeof(int), s));
    CK(down(tcur, dtcur, nb, s));
    if (want_rows) CK(down(rows, drows, rb, s));
    CK(cudaStreamSynchronize(s));
    return RAPT_OK;
}

// ------------------------------------------------------------------------------------------------
// small per-call kernels (host pointers)
// ------------------------------------------------------------------------------------------------

int rapt_b200_field_ops(const rapt_field_t *f, int arith, int64_t npt, const double *tpos,
                        double *B, double *E, double *unitb, double *magB, double *gradB, double *jacobianB,
                        double *curlb, double *curvature, double *dBdt, double *dbdt,
                        double *lengthscale, double *timescale)
{
    if (int rc = ensure_init()) return rc;
    if (int rc = check_field(f)) return rc;
    if (npt < 0 || (npt > 0 && !tpos)) return fail(RAPT_E_ARG, "field_ops: bad argument");
    if (npt == 0) return RAPT_OK;
    cudaStream_t s = 0;
    DevBuf dpos, o[12];
    CK(up(dpos, tpos, npt * 4 * sizeof(double), s));
    double *host[12] = {B, E, unitb, magB, gradB, jacobianB, curlb, curvature, dBdt, dbdt, lengthscale, timescale};
    const int width[12] = {3, 3, 3, 1, 3, 9, 3, 1, 1, 3, 1, 1};
    for (int k = 0; k < 12; k++) if (host[k]) CK(o[k].alloc(npt * width[k] * sizeof(double)));
    rapt::OpsArgs a;
    memset(&a, 0, sizeof a);
    RESOLVE(a.f, f);
    a.n = npt; a.tpos = dpos.as<double>();
    a.B = o[0].as<double>(); a.E = o[1].as<double>(); a.unitb = o[2].as<double>(); a.magB = o[3].as<double>();
    a.gradB = o[4].as<double>(); a.jac = o[5].as<double>(); a.curlb = o[6].as<double>(); a.curv = o[7].as<double>();
    a.dBdt = o[8].as<double>(); a.dbdt = o[9].as<double>(); a.lscale = o[10].as<double>(); a.tscale = o[11].as<double>();
    if (int rc = launch_any(f, arith == 1, UK_FIELD_OPS, &a, npt, 0, s)) return rc;
    g_launches++;
    for (int k = 0; k < 12; k++) if (host[k]) CK(down(host[k], o[k], npt * width[k] * sizeof(double), s));
    CK(cudaStreamSynchronize(s));
    return RAPT_OK;
}

int rapt_b200_gc_construct(const rapt_field_t *f, int arith, int64_t n, const double *t0, const double *x,
                           const double *y, const double *z, const double *v, const double *pa_deg,
                           const double *mass, double *ppar, double *mu)
{
    if (int rc = ensure_init()) return rc;
    if (int rc = check_field(f)) return rc;
    if (n < 0 || (n > 0 && (!t0 || !x || !y || !z || !v || !pa_deg || !mass || !ppar || !mu))) return fail(RAPT_E_ARG, "gc_construct: bad argument");
    if (n == 0) return RAPT_OK;
    cudaStream_t s = 0;
    const size_t nb = n * sizeof(double);
    DevBuf in[7], o0, o1;
    const double *h[7] = {t0, x, y, z, v, pa_deg, mass};
    for (int k = 0; k < 7; k++) CK(up(in[k], h[k], nb, s));
    CK(o0.alloc(nb)); CK(o1.alloc(nb));
    rapt::MiscArgs a;
    memset(&a, 0, sizeof a);
    RESOLVE(a.f, f);
    a.n = n; a.op = 0;
    a.a0 = in[0].as<double>(); a.a1 = in[1].as<double>(); a.a2 = in[2].as<double>(); a.a3 = in[3].as<double>();
    a.a4 = in[4].as<double>(); a.a5 = in[5].as<double>(); a.a6 = in[6].as<double>();
    a.o0 = o0.as<double>(); a.o1 = o1.as<double>();
    if (int rc = launch_any(f, arith == 1, UK_MISC, &a, n, 0, s)) return rc;
    g_launches++;
    CK(down(ppar, o0, nb, s)); CK(down(mu, o1, nb, s));
    CK(cudaStreamSynchronize(s));
    return RAPT_OK;
}

int rapt_b200_switch_p2g(const rapt_field_t *f, int arith, int64_t n, const double *prow7, const double *mass,
                         const double *charge, double *grow5, double *mu, double *v, int32_t *status)
{
    if (int rc = ensure_init()) return rc;
    if (int rc = check_field(f)) return rc;
    if (n < 0 || (n > 0 && (!prow7 || !mass || !charge || !grow5 || !mu || !v || !status))) return fail(RAPT_E_ARG, "switch_p2g: bad argument");
    if (n == 0) return RAPT_OK;
    cudaStream_t s = 0;
    const size_t nb = n * sizeof(double);
    DevBuf dp, dm, dq, og, omu, ov, ost;
    CK(up(dp, prow7, 7 * nb, s)); CK(up(dm, mass, nb, s)); CK(up(dq, charge, nb, s));
    CK(og.alloc(5 * nb)); CK(omu.alloc(nb)); CK(ov.alloc(nb)); CK(ost.alloc(n * sizeof(int)));
    rapt::MiscArgs a;
    memset(&a, 0, sizeof a);
    RESOLVE(a.f, f);
    a.n = n; a.op = 1;
    a.a0 = dp.as<double>(); a.a1 = dm.as<double>(); a.a2 = dq.as<double>();
    a.o0 = og.as<double>(); a.o1 = omu.as<double>(); a.o2 = ov.as<double>(); a.io = ost.as<int>();
    if (int rc = launch_any(f, arith == 1, UK_MISC, &a, n, 0, s)) return rc;
    g_launches++;
    CK(down(grow5, og, 5 * nb, s)); CK(down(mu, omu, nb, s)); CK(down(v, ov, nb, s)); CK(down(status, ost, n * sizeof(int), s));
    CK(cudaStreamSynchronize(s));
    return RAPT_OK;
}

int rapt_b200_switch_g2p(const rapt_field_t *f, int arith, int64_t n, const double *grow5, const double *mu,
                         const double *mass, const double *charge, double t_eval, double *prow7)
{
    if (int rc = ensure_init()) return rc;
    if (int rc = check_field(f)) return rc;
    if (n < 0 || (n > 0 && (!grow5 || !mu || !mass || !charge || !prow7))) return fail(RAPT_E_ARG, "switch_g2p: bad argument");
    if (n == 0) return RAPT_OK;
    cudaStream_t s = 0;
    const size_t nb = n * sizeof(double);
    DevBuf dg, dmu, dm, dq, op;
    CK(up(dg, grow5, 5 * nb, s)); CK(up(dmu, mu, nb, s)); CK(up(dm, mass, nb, s)); CK(up(dq, charge, nb, s));
    CK(op.alloc(7 * nb));
    rapt::MiscArgs a;
    memset(&a, 0, sizeof a);
    RESOLVE(a.f, f);
    a.n = n; a.op = 2; a.t_eval = t_eval;
    a.a0 = dg.as<double>(); a.a1 = dmu.as<double>(); a.a2 = dm.as<double>(); a.a3 = dq.as<double>();
    a.o0 = op.as<double>();
    if (int rc = launch_any(f, arith == 1, UK_MISC, &a, n, 0, s)) return rc;
    g_launches++;
    CK(down(prow7, op, 7 * nb, s));
    CK(cudaStreamSynchronize(s));
    return RAPT_OK;
}

int rapt_b200_isadiabatic(const rapt_field_t *f, const rapt_params_t *p, int mode, int64_t n, const double *rows,
                          int64_t row_stride, const double *mu, const double *mass, const double *charge, int32_t *out)
{
    if (int rc = ensure_init()) return rc;
    if (int rc = check_field(f)) return rc;
    if (!p || n < 0 || (n > 0 && (!rows || !mass || !charge || !out || (mode == 1 && !mu)))) return fail(RAPT_E_ARG, "isadiabatic: bad argument");
    if (row_stride < (mode == 0 ? 7 : 5)) return fail(RAPT_E_ARG, "isadiabatic: row_stride too small");
    if (n == 0) return RAPT_OK;
    cudaStream_t s = 0;
    const size_t nb = n * sizeof(double);
    DevBuf dr, dmu, dm, dq, oo;
    CK(up(dr, rows, row_stride * nb, s)); if (mu) CK(up(dmu, mu, nb, s));
    CK(up(dm, mass, nb, s)); CK(up(dq, charge, nb, s)); CK(oo.alloc(n * sizeof(int)));
    rapt::MiscArgs a;
    memset(&a, 0, sizeof a);
    RESOLVE(a.f, f); memcpy(&a.p, p, sizeof a.p);
    a.n = n; a.op = 3; a.mode = mode; a.stride = row_stride;
    a.a0 = dr.as<double>(); a.a1 = dmu.as<double>(); a.a2 = dm.as<double>(); a.a3 = dq.as<double>(); a.io = oo.as<int>();
    if (int rc = launch_any(f, p->arith == 1, UK_MISC, &a, n, 0, s)) return rc;
    g_launches++;
    CK(down(out, oo, n * sizeof(int), s));
    CK(cudaStreamSynchronize(s));
    return RAPT_OK;
}

static int bounce_impl(const rapt_field_t *f, int arith, double fieldlineresolution, int64_t n,
                       const double *t, const double *x, const double *y, const double *z, const double *ppar,
                       const double *mu, const double *mass,
                       double *Bm, double *v, double *ds, int32_t *npts, int64_t max_pts, double *curve, double *period,
                       int quadrature = 0);

int rapt_b200_bounce_setup(const rapt_field_t *f, int arith, double fieldlineresolution, int64_t n,
                           const double *t, const double *x, const double *y, const double *z, const double *ppar,
                           const double *mu, const double *mass,
                           double *Bm, double *v, double *ds, int32_t *npts, int64_t max_pts, double *curve)
{
    if (!curve) return fail(RAPT_E_ARG, "bounce_setup: curve buffer required");
    return bounce_impl(f, arith, fieldlineresolution, n, t, x, y, z, ppar, mu, mass, Bm, v, ds, npts, max_pts, curve, nullptr);
}

int rapt_b200_bounce_period(const rapt_field_t *f, int arith, int quadrature, double fieldlineresolution, int64_t n,
                            const double *t, const double *x, const double *y, const double *z, const double *ppar,
                            const double *mu, const double *mass, double *period, int32_t *npts)
{
    if (!period || !mu || !ppar || !mass) return fail(RAPT_E_ARG, "bounce_period: null argument");
    if (n <= 0) return n < 0 ? fail(RAPT_E_ARG, "n < 0") : RAPT_OK;
    std::vector<double> Bm((size_t)n), v((size_t)n), ds((size_t)n);
    std::vector<int32_t> np_local((size_t)n);
    int32_t *np_out = npts ? npts : np_local.data();
    for (int64_t max_pts = 128; max_pts <= 8192; max_pts *= 4) {
        int rc = bounce_impl(f, arith, fieldlineresolution, n, t, x, y, z, ppar, mu, mass, Bm.data(), v.data(), ds.data(),
                             np_out, max_pts, nullptr, period, quadrature);
        if (rc) return rc;
        int32_t mx = 0;
        for (int64_t i = 0; i < n; i++) mx = std::max(mx, np_out[i]);
        if (mx <= max_pts) return RAPT_OK;
    }
    return fail(RAPT_E_ARG, "bounce_period: a field line needs more than 8192 points");
}

static int bounce_impl(const rapt_field_t *f, int arith, double fieldlineresolution, int64_t n,
                       const double *t, const double *x, const double *y, const double *z, const double *ppar,
                       const double *mu, const double *mass,
                       double *Bm, double *v, double *ds, int32_t *npts, int64_t max_pts, double *curve, double *period,
                       int quadrature)
{
    if (int rc = ensure_init()) return rc;
    if (int rc = check_field(f)) return rc;
    if (n < 0 || max_pts < 3 || (n > 0 && (!t || !x || !y || !z || !Bm || !ds || !npts)))
        return fail(RAPT_E_ARG, "bounce_setup: bad argument");
    if (mu && (!ppar || !mass || !v)) return fail(RAPT_E_ARG, "bounce_setup: ppar, mass, v required with mu");
    if (n == 0) return RAPT_OK;
    cudaStream_t s = 0;
    const size_t nb = n * sizeof(double);
    DevBuf in[7], oBm, ov, ods, onp, ocv, scr, oper;
    const double *h[7] = {t, x, y, z, ppar, mu, mass};
    for (int k = 0; k < 7; k++) if (h[k]) CK(up(in[k], h[k], nb, s));
    if (mu) CK(oBm.alloc(nb)); else CK(up(oBm, Bm, nb, s));      // mu == NULL: Bm is an input
    CK(ov.alloc(nb)); CK(ods.alloc(nb)); CK(onp.alloc(n * sizeof(int)));
    CK(ocv.alloc((size_t)n * max_pts * 5 * sizeof(double))); CK(scr.alloc((size_t)n * max_pts * 4 * sizeof(double)));
    rapt::BounceArgs a;
    memset(&a, 0, sizeof a);
    RESOLVE(a.f, f);
    a.flres = fieldlineresolution; a.n = n; a.max_pts = max_pts;
    a.t = in[0].as<double>(); a.x = in[1].as<double>(); a.y = in[2].as<double>(); a.z = in[3].as<double>();
    a.ppar = in[4].as<double>(); a.mu = in[5].as<double>(); a.mass = in[6].as<double>();
    a.Bm = oBm.as<double>(); a.v = ov.as<double>(); a.ds = ods.as<double>(); a.npts = onp.as<int>();
    a.curve = ocv.as<double>(); a.scratch = scr.as<double>();
    if (period) { CK(oper.alloc(nb)); a.period = oper.as<double>(); }
    a.quadrature = quadrature;
    if (int rc = launch_any(f, arith == 1, UK_BOUNCE, &a, n, 0, s)) return rc;
    g_launches++;
    CK(down(Bm, oBm, nb, s)); if (v) CK(down(v, ov, nb, s)); CK(down(ds, ods, nb, s)); CK(down(npts, onp, n * sizeof(int), s));
    if (curve) CK(down(curve, ocv, (size_t)n * max_pts * 5 * sizeof(double), s));
    if (period) CK(down(period, oper, nb, s));
    CK(cudaStreamSynchronize(s));
    return RAPT_OK;
}

// ------------------------------------------------------------------------------------------------
// BounceCenter.advance and the flutils pieces behind it (rapt_bc.cuh).  One lane per tracer up to
// 148 SMs x 4 blocks x 64 threads; every lane owns a scratch curve of max_pts points, grown on overflow.
// ------------------------------------------------------------------------------------------------
static int bc_run(const rapt_field_t *f, int arith, rapt::BCArgs &a, int64_t n, cudaStream_t s, bool any_time_dependence = false)
{
    if (f->kind == RAPT_FIELD_GRID)
        return fail(RAPT_E_UNSUPPORTED, "bounce centre: analytic fields only (built-in or NVRTC user fields)");
    // flutils.eye traces the line at the time of its start point, whatever the field (GuidingCenter.geteye); the
    // bounce-centre tracer itself is for static fields only (BounceCenter.py:104-105)
    if (!any_time_dependence && !f->is_static) return fail(RAPT_E_ARG, "BounceCenter does not work with nonstatic fields or electric fields.");
    // 128 registers per thread: up to 8 blocks of 64 threads (16 warps) are resident per SM
    static const int blocks_per_sm = getenv("RAPT_B200_BC_BLOCKS") ? std::max(1, atoi(getenv("RAPT_B200_BC_BLOCKS"))) : 8;
    const long long lanes_max = (long long)g_sms * blocks_per_sm * 64;
    const long long lanes = std::min<long long>(n, lanes_max);
    const int grid = (int)((lanes + 63) / 64);
    DevBuf cv, bw;
    CK(cv.alloc((size_t)grid * 64 * a.max_pts * 5 * sizeof(double)));
    CK(bw.alloc((size_t)grid * 64 * a.max_pts * 4 * sizeof(double)));
    a.curve = cv.as<double>(); a.scratch = bw.as<double>();
    if (int rc = launch_any(f, arith == 1, UK_BC, &a, n, grid, s)) return rc;
    g_launches++;
    CK(cudaStreamSynchronize(s));
    return RAPT_OK;
}

int rapt_b200_bounce_center_advance(const rapt_field_t *f, int arith, int quadrature, int64_t n,
                                    double *t, double *x, double *y, double *z,
                                    const double *mu, const double *v, const double *mass, const double *charge,
                                    const double *dt_in, double bctimestep, double delta,
                                    double rtol, double atol, double fieldlineresolution, double eyegradientstep,
                                    int64_t store_every, int64_t max_rows, double *rows,
                                    int32_t *nrows, int32_t *nstored, int32_t *counters, int32_t *status, double *dt_out)
{
    if (int rc = ensure_init()) return rc;
    if (int rc = check_field(f)) return rc;
    if (n < 0 || (n > 0 && (!t || !x || !y || !z || !mu || !v || !mass || !charge || !nrows || !nstored || !counters || !status)))
        return fail(RAPT_E_ARG, "bounce_center_advance: null argument");
    if (!(fieldlineresolution > 0) || !(eyegradientstep > 0) || !(rtol > 0) || !(atol >= 0))
        return fail(RAPT_E_ARG, "bounce_center_advance: bad parameter");
    if (!dt_in && !(bctimestep > 0)) return fail(RAPT_E_ARG, "bounce_center_advance: BCtimestep must be > 0");
    if (f->kind == RAPT_FIELD_GRID)
        return fail(RAPT_E_UNSUPPORTED, "bounce centre: analytic fields only (built-in or NVRTC user fields)");
    if (!f->is_static) return fail(RAPT_E_ARG, "BounceCenter does not work with nonstatic fields or electric fields.");
    if (n == 0) return RAPT_OK;
    cudaStream_t s = 0;
    const size_t nb = (size_t)n * sizeof(double), ni = (size_t)n * sizeof(int);
    const bool want_rows = rows && max_rows > 0 && store_every > 0;
    // Work ordering: one trace of the field line through every start point (the kernel of the bounce-period set-up) gives
    // its number of points; tracers are handed to the lanes longest line first, so the lanes of a warp follow lines of
    // similar length through the five traces of every right-hand side (14.6 of 32 lanes were active without it).
    std::vector<int> order;
    if (n > 64 && !(getenv("RAPT_B200_BC_SORT") && atoi(getenv("RAPT_B200_BC_SORT")) == 0)) {
        std::vector<double> hBm((size_t)n), hds((size_t)n);
        std::vector<int32_t> hnp((size_t)n);
        for (int64_t i = 0; i < n; i++) {
            const double vc = v[i] / 299792458.0, g2 = 1.0 / (1 - vc * vc);
            hBm[i] = mass[i] * g2 * v[i] * v[i] / (2 * mu[i]);            // BounceCenter.py:226-227
        }
        const int64_t chunk = 1 << 18;
        for (int64_t o = 0; o < n; o += chunk) {
            const int64_t m = std::min<int64_t>(chunk, n - o);
            if (int rc = bounce_impl(f, arith, fieldlineresolution, m, t + o, x + o, y + o, z + o, nullptr, nullptr, nullptr,
                                     hBm.data() + o, nullptr, hds.data() + o, hnp.data() + o, 256, nullptr, nullptr)) return rc;
        }
        order.resize((size_t)n);
        for (int64_t i = 0; i < n; i++) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int a_, int b_) { return hnp[a_] > hnp[b_]; });
    }
    for (int64_t max_pts = 256; max_pts <= 16384; max_pts *= 4) {
        DevBuf st[4], in[4], ddt, dodt, drows, dnr, dns, dcnt, dstat, dord;
        if (!order.empty()) CK(up(dord, order.data(), ni, s));
        double *hs[4] = {t, x, y, z};
        const double *hi[4] = {mu, v, mass, charge};
        for (int k = 0; k < 4; k++) { CK(up(st[k], hs[k], nb, s)); CK(up(in[k], hi[k], nb, s)); }
        if (dt_in) CK(up(ddt, dt_in, nb, s));
        CK(dodt.alloc(nb)); CK(dnr.alloc(ni)); CK(dns.alloc(ni)); CK(dcnt.alloc(4 * ni)); CK(dstat.alloc(ni));
        if (want_rows) CK(drows.alloc((size_t)n * max_rows * 4 * sizeof(double)));
        rapt::BCArgs a;
        memset(&a, 0, sizeof a);
        RESOLVE(a.f, f);
        a.op = 0; a.quadrature = quadrature; a.rtol = rtol; a.atol = atol; a.flres = fieldlineresolution;
        a.eyestep = eyegradientstep; a.bctimestep = bctimestep; a.delta = delta;
        a.n = n; a.max_pts = max_pts; a.store_every = want_rows ? store_every : 0; a.max_rows = want_rows ? max_rows : 0;
        a.t = st[0].as<double>(); a.x = st[1].as<double>(); a.y = st[2].as<double>(); a.z = st[3].as<double>();
        a.mu = in[0].as<double>(); a.v = in[1].as<double>(); a.mass = in[2].as<double>(); a.charge = in[3].as<double>();
        a.dtin = dt_in ? ddt.as<double>() : nullptr; a.dt_out = dodt.as<double>();
        a.rows = want_rows ? drows.as<double>() : nullptr;
        a.nrows = dnr.as<int>(); a.nstored = dns.as<int>(); a.counters = dcnt.as<int>(); a.status = dstat.as<int>();
        a.order = order.empty() ? nullptr : dord.as<int>();
        if (int rc = bc_run(f, arith, a, n, s)) return rc;
        CK(down(status, dstat, ni, s));
        CK(cudaStreamSynchronize(s));
        bool overflow = false;
        for (int64_t i = 0; i < n; i++) if (status[i] == RAPT_ST_ROWCAP) { overflow = true; break; }
        if (overflow && max_pts < 16384) continue;          // a field line needed more points: rerun with larger scratch
        for (int k = 0; k < 4; k++) CK(down(hs[k], st[k], nb, s));
        CK(down(nrows, dnr, ni, s)); CK(down(nstored, dns, ni, s)); CK(down(counters, dcnt, 4 * ni, s));
        if (dt_out) CK(down(dt_out, dodt, nb, s));
        if (want_rows) CK(down(rows, drows, (size_t)n * max_rows * 4 * sizeof(double), s));
        CK(cudaStreamSynchronize(s));
        return RAPT_OK;
    }
    return RAPT_OK;
}

int rapt_b200_bounce_center_terms(const rapt_field_t *f, int arith, int quadrature, int64_t n,
                                  const double *t, const double *x, const double *y, const double *z, const double *Bm,
                                  const double *v, const double *mass, const double *charge,
                                  double fieldlineresolution, double eyegradientstep, double *out, int32_t *status)
{
    if (int rc = ensure_init()) return rc;
    if (int rc = check_field(f)) return rc;
    if (n < 0 || (n > 0 && (!t || !x || !y || !z || !Bm || !out || !status)))
        return fail(RAPT_E_ARG, "bounce_center_terms: null argument");
    if (!(fieldlineresolution > 0) || !(eyegradientstep > 0)) return fail(RAPT_E_ARG, "bounce_center_terms: bad parameter");
    if (n == 0) return RAPT_OK;
    cudaStream_t s = 0;
    const size_t nb = (size_t)n * sizeof(double), ni = (size_t)n * sizeof(int);
    std::vector<double> nanv;
    if (!v || !mass || !charge) nanv.assign((size_t)n, std::nan(""));
    for (int64_t max_pts = 256; max_pts <= 16384; max_pts *= 4) {
        DevBuf in[8], dout, dstat;
        const double *hi[8] = {t, x, y, z, Bm, v ? v : nanv.data(), mass ? mass : nanv.data(), charge ? charge : nanv.data()};
        for (int k = 0; k < 8; k++) CK(up(in[k], hi[k], nb, s));
        CK(dout.alloc(8 * nb)); CK(dstat.alloc(ni));
        rapt::BCArgs a;
        memset(&a, 0, sizeof a);
        RESOLVE(a.f, f);
        a.op = 1; a.quadrature = quadrature; a.flres = fieldlineresolution; a.eyestep = eyegradientstep;
        a.n = n; a.max_pts = max_pts;
        a.t = in[0].as<double>(); a.x = in[1].as<double>(); a.y = in[2].as<double>(); a.z = in[3].as<double>();
        a.Bm = in[4].as<double>(); a.v = in[5].as<double>(); a.mass = in[6].as<double>(); a.charge = in[7].as<double>();
        a.out = dout.as<double>(); a.status = dstat.as<int>();
        if (int rc = bc_run(f, arith, a, n, s)) return rc;
        CK(down(status, dstat, ni, s));
        CK(cudaStreamSynchronize(s));
        bool overflow = false;
        for (int64_t i = 0; i < n; i++) if (status[i] == RAPT_ST_ROWCAP) { overflow = true; break; }
        if (overflow && max_pts < 16384) continue;
        CK(down(out, dout, 8 * nb, s));
        CK(cudaStreamSynchronize(s));
        return RAPT_OK;
    }
    return RAPT_OK;
}

int rapt_b200_second_invariant(const rapt_field_t *f, int arith, int64_t n,
                               const double *t, const double *x, const double *y, const double *z, const double *Bm,
                               double fieldlineresolution, double *I, int32_t *status)
{
    if (int rc = ensure_init()) return rc;
    if (int rc = check_field(f)) return rc;
    if (n < 0 || (n > 0 && (!t || !x || !y || !z || !Bm || !I || !status))) return fail(RAPT_E_ARG, "second_invariant: null argument");
    if (!(fieldlineresolution > 0)) return fail(RAPT_E_ARG, "second_invariant: bad parameter");
    if (f->kind == RAPT_FIELD_GRID) return fail(RAPT_E_UNSUPPORTED, "second_invariant: analytic fields only");
    if (n == 0) return RAPT_OK;
    cudaStream_t s = 0;
    const size_t nb = (size_t)n * sizeof(double), ni = (size_t)n * sizeof(int);
    for (int64_t max_pts = 256; max_pts <= 16384; max_pts *= 4) {
        DevBuf in[5], dout, dstat;
        const double *hi[5] = {t, x, y, z, Bm};
        for (int k = 0; k < 5; k++) CK(up(in[k], hi[k], nb, s));
        CK(dout.alloc(nb)); CK(dstat.alloc(ni));
        rapt::BCArgs a;
        memset(&a, 0, sizeof a);
        RESOLVE(a.f, f);
        a.op = 2; a.quadrature = RAPT_QUAD_QUADPACK; a.flres = fieldlineresolution; a.eyestep = 1.0;
        a.n = n; a.max_pts = max_pts;
        a.t = in[0].as<double>(); a.x = in[1].as<double>(); a.y = in[2].as<double>(); a.z = in[3].as<double>();
        a.Bm = in[4].as<double>();
        a.v = a.Bm; a.mass = a.Bm; a.charge = a.Bm;        // read but unused by op 2
        a.out = dout.as<double>(); a.status = dstat.as<int>();
        if (int rc = bc_run(f, arith, a, n, s, /*any_time_dependence=*/true)) return rc;
        CK(down(status, dstat, ni, s));
        CK(cudaStreamSynchronize(s));
        bool overflow = false;
        for (int64_t i = 0; i < n; i++) if (status[i] == RAPT_ST_ROWCAP) { overflow = true; break; }
        if (overflow && max_pts < 16384) continue;
        CK(down(I, dout, nb, s));
        CK(cudaStreamSynchronize(s));
        return RAPT_OK;
    }
    return RAPT_OK;
}

// ------------------------------------------------------------------------------------------------
// Adaptive.__init__ + Adaptive.advance for an ensemble: epochs of
//   [particle kernel over the particle-mode list | guiding-centre kernel over the GC-mode list]
//   -> switch + compaction kernel
// until no tracer has t < delta.  The two advance kernels of an epoch are independent and run on two
// streams.
// ------------------------------------------------------------------------------------------------
int rapt_b200_adaptive_advance(const rapt_field_t *f, const rapt_params_t *p, int64_t n,
                               const double *x, const double *y, const double *z,
                               const double *vx, const double *vy, const double *vz,
                               const double *t0, const double *mass, const double *charge,
                               double gc_dt, double delta, int64_t store_every, int64_t max_rows, double *rows,
                               int32_t *nstored, int32_t *nseg, int32_t *mode_out, double *fin,
                               int32_t *counters, int32_t *status, int32_t *epochs_out)
{
    if (int rc = ensure_init()) return rc;
    if (int rc = check_field(f)) return rc;
    if (!p || n < 0 || (n > 0 && (!x || !y || !z || !vx || !vy || !vz || !t0 || !mass || !charge)))
        return fail(RAPT_E_ARG, "adaptive_advance: null argument");
    if (!(gc_dt > 0)) return fail(RAPT_E_ARG, "adaptive_advance: gc_dt (params['GCtimestep']) must be > 0 for ensembles");
    if (n == 0) { if (epochs_out) *epochs_out = 0; return RAPT_OK; }
    const bool strict = p->arith == 1;
    const bool want_rows = rows && max_rows > 0;
    const size_t nb = (size_t)n * sizeof(double), ni = (size_t)n * sizeof(int);
    cudaStream_t s0 = 0, s1 = nullptr, s2 = nullptr;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t a, b; ~StreamGuard() { cudaStreamDestroy(a); cudaStreamDestroy(b); } } guard{s1, s2};
    DevBuf in[9];
    const double *h[9] = {x, y, z, vx, vy, vz, t0, mass, charge};
    for (int k = 0; k < 9; k++) CK(up(in[k], h[k], nb, s0));
    DevBuf ps[7], gs[7], dmode, dst, dnseg, dtag, dnst, dtvar, drem, dtcur, drows, dlp, dlg, dcounts, dcnt, ddtg, ddtp;
    DevBuf dsts, dsx, dsdt, dsrow, dcntg;
    CK(dcntg.alloc(4 * ni)); CK(cudaMemsetAsync(dcntg.p, 0, 4 * ni, s0));
    AdaptiveStats stt;
    cudaEvent_t ev[6];
    for (auto &e : ev) CK(cudaEventCreate(&e));
    struct EventGuard { cudaEvent_t *e; ~EventGuard() { for (int k = 0; k < 6; k++) cudaEventDestroy(e[k]); } } eguard{ev};
    CK(cudaEventRecord(ev[4], s0));
    for (int k = 0; k < 7; k++) { CK(ps[k].alloc(nb)); CK(gs[k].alloc(nb)); CK(cudaMemsetAsync(gs[k].p, 0, nb, s0)); CK(cudaMemsetAsync(ps[k].p, 0, nb, s0)); }
    CK(dmode.alloc(ni)); CK(dst.alloc(ni)); CK(dnseg.alloc(ni)); CK(dtag.alloc(ni)); CK(dnst.alloc(ni));
    CK(dtvar.alloc(nb)); CK(drem.alloc(nb)); CK(dtcur.alloc(nb));
    CK(drows.alloc(want_rows ? (size_t)n * max_rows * 8 * sizeof(double) : 0));
    CK(dlp.alloc(ni)); CK(dlg.alloc(ni)); CK(dcounts.alloc(2 * sizeof(int))); CK(dcnt.alloc(4 * ni));
    CK(ddtg.alloc(nb)); CK(ddtp.alloc(nb));
    CK(dsts.alloc(nb)); CK(dsx.alloc(nb)); CK(dsdt.alloc(nb)); CK(dsrow.alloc(ni));
    CK(cudaMemsetAsync(dsts.p, 0, nb, s0)); CK(cudaMemsetAsync(dsx.p, 0, nb, s0)); CK(cudaMemsetAsync(dsdt.p, 0, nb, s0)); CK(cudaMemsetAsync(dsrow.p, 0, ni, s0));
    CK(cudaMemsetAsync(dcnt.p, 0, 4 * ni, s0)); CK(cudaMemsetAsync(dcounts.p, 0, 2 * sizeof(int), s0));
    {   // per-tracer GC output step (uniform: params["GCtimestep"])
        std::vector<double> hd((size_t)n, gc_dt);
        CK(cudaMemcpyAsync(ddtg.p, hd.data(), nb, cudaMemcpyHostToDevice, s0));
        CK(cudaStreamSynchronize(s0));
    }
    rapt::AdaptArgs sw;
    memset(&sw, 0, sizeof sw);
    RESOLVE(sw.f, f); memcpy(&sw.p, p, sizeof sw.p);
    sw.n = n; sw.delta = delta;
    sw.x0 = in[0].as<double>(); sw.y0 = in[1].as<double>(); sw.z0 = in[2].as<double>();
    sw.vx0 = in[3].as<double>(); sw.vy0 = in[4].as<double>(); sw.vz0 = in[5].as<double>(); sw.t0 = in[6].as<double>();
    sw.mass = in[7].as<double>(); sw.charge = in[8].as<double>();
    sw.pt = ps[0].as<double>(); sw.px = ps[1].as<double>(); sw.py = ps[2].as<double>(); sw.pz = ps[3].as<double>();
    sw.ppx = ps[4].as<double>(); sw.ppy = ps[5].as<double>(); sw.ppz = ps[6].as<double>();
    sw.gt = gs[0].as<double>(); sw.gx = gs[1].as<double>(); sw.gy = gs[2].as<double>(); sw.gz = gs[3].as<double>();
    sw.gpp = gs[4].as<double>(); sw.mu = gs[5].as<double>(); sw.v = gs[6].as<double>();
    sw.mode = dmode.as<int>(); sw.status = dst.as<int>(); sw.nseg = dnseg.as<int>(); sw.segtag = dtag.as<int>();
    sw.nstored = dnst.as<int>(); sw.tvar = dtvar.as<double>(); sw.rem = drem.as<double>(); sw.tcur = dtcur.as<double>();
    sw.max_rows = want_rows ? max_rows : 0; sw.rows = want_rows ? drows.as<double>() : nullptr;
    sw.listP = dlp.as<int>(); sw.listG = dlg.as<int>(); sw.counts = dcounts.as<int>();
    sw.seg_tstop = dsts.as<double>(); sw.seg_x = dsx.as<double>(); sw.seg_dt = dsdt.as<double>(); sw.seg_row = dsrow.as<int>();
    sw.first = 1;
    if (int rc = launch_any(f, strict, UK_ADAPT, &sw, n, 0, s0)) return rc;
    g_launches++;
    sw.first = 0;

    rapt_params_t pc = *p;
    pc.check_adiabaticity = 1;
    const int gp = grid_for(1 << 30, FLAVOUR(strict, particle_blocks_per_sm, !strict && f->is_static && !p->enforce_equatorial && f->kind != RAPT_FIELD_USER));
    const int gg = grid_for(1 << 30, FLAVOUR(strict, gc_blocks_per_sm));
    // Optional time-sliced epochs (RAPT_B200_ADAPTIVE_SLICES=k): every launch advances its tracers at most to
    // slice_end, so a tracer that switches mode early waits one slice instead of the longest segment of the
    // ensemble.  Slices only interrupt at row boundaries and resume bit-identically (tests run with 16).
    // Measured on config 4 it does NOT pay: 1 M Speiser tracers 3.7 s unsliced vs 5.1 s with 8-32 slices,
    // 65,536 tracers 1.6 s vs 2.2-2.4 s (profiles/r1_other_configs.md) -- the relaunch of every tracer per slice
    // costs more than the idle lanes it removes -- so the default is one slice.
    double tmin = t0[0];
    for (int64_t i = 1; i < n; i++) tmin = std::min(tmin, t0[i]);
    const char *env_sl = getenv("RAPT_B200_ADAPTIVE_SLICES");
    const int nslices = env_sl ? std::max(1, atoi(env_sl)) : 1;
    const double slice = (delta > 0 ? delta : 1.0) / nslices;
    int epochs = 0;
    for (;; epochs++) {
        const double slice_end = tmin + slice * (epochs + 1);
        int cnt[2] = {0, 0};
        CK(cudaMemcpyAsync(cnt, dcounts.p, sizeof cnt, cudaMemcpyDeviceToHost, s0));
        CK(cudaStreamSynchronize(s0));
        if (cnt[0] == 0 && cnt[1] == 0) break;
        if (epochs > 100000) return fail(RAPT_E_CUDA, "adaptive_advance: epoch limit");
        if (getenv("RAPT_B200_TRACE")) fprintf(stderr, "[rapt_b200] adaptive epoch %d: %d particle-mode, %d guiding-centre-mode tracers\n", epochs, cnt[0], cnt[1]);
        CK(cudaMemsetAsync(dcounts.p, 0, 2 * sizeof(int), s0));
        CK(cudaStreamSynchronize(s0));
        if (cnt[0] > 0) {
            rapt::AdvArgs a;
            if (int rc = fill_common(a, f, &pc)) return rc;
            a.nwork = cnt[0]; a.order = sw.listP; a.queue = next_queue(s1);
            a.t = sw.pt; a.s1 = sw.px; a.s2 = sw.py; a.s3 = sw.pz; a.s4 = sw.ppx; a.s5 = sw.ppy; a.s6 = sw.ppz;
            a.mass = sw.mass; a.charge = sw.charge; a.delta = 0; a.delta_arr = sw.rem;
            a.store_every = want_rows ? std::max<int64_t>(store_every, 1) : 0; a.max_rows = sw.max_rows; a.rows = sw.rows;
            a.nstored = sw.nstored; a.nrows = nullptr; a.counters = dcnt.as<int>(); a.status = sw.status;
            a.tcur = sw.tcur; a.dt_out = ddtp.as<double>(); a.segtag = sw.segtag; a.append = 1;
            a.seg_tstop = sw.seg_tstop; a.seg_x = sw.seg_x; a.seg_dt = sw.seg_dt; a.seg_row = sw.seg_row; a.slice_end = slice_end;
            CK(cudaEventRecord(ev[0], s1));
            if (int rc = launch_any(f, strict, UK_PARTICLE, &a, cnt[0], std::min(gp, (cnt[0] + 127) / 128), s1)) return rc;
            CK(cudaEventRecord(ev[1], s1));
            g_launches++; stt.launches_p++; stt.tracer_launches_p += cnt[0];
        }
        if (cnt[1] > 0) {
            rapt::AdvArgs a;
            if (int rc = fill_common(a, f, &pc)) return rc;
            a.nwork = cnt[1]; a.order = sw.listG; a.queue = next_queue(s2);
            a.t = sw.gt; a.s1 = sw.gx; a.s2 = sw.gy; a.s3 = sw.gz; a.s4 = sw.gpp;
            a.mass = sw.mass; a.charge = sw.charge; a.mu = sw.mu; a.v = sw.v; a.dtin = ddtg.as<double>();
            a.delta = 0; a.delta_arr = sw.rem; a.eom = RAPT_EOM_TAOCHANBRIZARD;
            a.store_every = want_rows ? std::max<int64_t>(store_every, 1) : 0; a.max_rows = sw.max_rows; a.rows = sw.rows;
            a.nstored = sw.nstored; a.nrows = nullptr; a.counters = dcntg.as<int>(); a.status = sw.status;
            a.tcur = sw.tcur; a.segtag = sw.segtag; a.append = 1;
            a.seg_tstop = sw.seg_tstop; a.seg_x = sw.seg_x; a.seg_dt = sw.seg_dt; a.seg_row = sw.seg_row; a.slice_end = slice_end;
            CK(cudaEventRecord(ev[2], s2));
            if (int rc = launch_any(f, strict, UK_GC, &a, cnt[1], std::min(gg, (cnt[1] + 127) / 128), s2)) return rc;
            CK(cudaEventRecord(ev[3], s2));
            g_launches++; stt.launches_g++; stt.tracer_launches_g += cnt[1];
        }
        CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2));
        float ms = 0;
        float msp = 0, msg = 0;
        if (cnt[0] > 0 && cudaEventElapsedTime(&msp, ev[0], ev[1]) == cudaSuccess) stt.ms_particle += msp;
        if (cnt[1] > 0 && cudaEventElapsedTime(&msg, ev[2], ev[3]) == cudaSuccess) stt.ms_gc += msg;
        stt.per_epoch.insert(stt.per_epoch.end(), {(float)cnt[0], (float)cnt[1], msp, msg});
        CK(cudaEventRecord(ev[0], s0));
        if (int rc = launch_any(f, strict, UK_ADAPT, &sw, n, 0, s0)) return rc;
        CK(cudaEventRecord(ev[1], s0));
        CK(cudaEventSynchronize(ev[1]));
        if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) stt.ms_switch += ms;
        g_launches++;
    }
    CK(cudaEventRecord(ev[5], s0));
    if (epochs_out) *epochs_out = epochs;
    // results
    std::vector<int> hmode((size_t)n), hcg(4 * (size_t)n), hcp(4 * (size_t)n);
    CK(cudaMemcpyAsync(hmode.data(), dmode.p, ni, cudaMemcpyDeviceToHost, s0));
    CK(down(nstored, dnst, ni, s0)); CK(down(nseg, dnseg, ni, s0)); CK(down(status, dst, ni, s0));
    CK(cudaMemcpyAsync(hcp.data(), dcnt.p, 4 * ni, cudaMemcpyDeviceToHost, s0));
    CK(cudaMemcpyAsync(hcg.data(), dcntg.p, 4 * ni, cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));
    {   // per-mode totals for the roofline of the executed mix; the ABI's counters are the sum, as the reference counts
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ev[4], ev[5]) == cudaSuccess) stt.ms_epochs = ms;
        for (int64_t i = 0; i < n; i++) {
            const int *cp = &hcp[4 * i], *cg = &hcg[4 * i];
            stt.steps_p += cp[1]; stt.accepted_p += cp[2]; stt.calls_p += (cp[0] - 11LL * cp[1] - cp[2]) / 2;
            stt.steps_g += cg[1]; stt.calls_g += (cg[0] - 6LL * cg[1]) / 2;
            if (counters) for (int k = 0; k < 4; k++) counters[4 * i + k] = cp[k] + cg[k];
        }
        stt.epochs = epochs;
        g_adaptive_stats = stt;
    }
    if (want_rows) CK(down(rows, drows, (size_t)n * max_rows * 8 * sizeof(double), s0));
    std::vector<double> hp[7], hg[7];
    if (fin) {
        for (int k = 0; k < 7; k++) {
            hp[k].resize(n); hg[k].resize(n);
            CK(cudaMemcpyAsync(hp[k].data(), ps[k].p, nb, cudaMemcpyDeviceToHost, s0));
            CK(cudaMemcpyAsync(hg[k].data(), gs[k].p, nb, cudaMemcpyDeviceToHost, s0));
        }
    }
    CK(cudaStreamSynchronize(s0));
    for (int64_t i = 0; i < n; i++) {
        if (mode_out) mode_out[i] = hmode[i];
        if (fin) {
            double *o = fin + 8 * i;
            if (hmode[i] == 0) { for (int k = 0; k < 7; k++) o[k] = hp[k][i]; o[7] = 0; }
            else { for (int k = 0; k < 5; k++) o[k] = hg[k][i]; o[5] = hg[5][i]; o[6] = hg[6][i]; o[7] = 1; }
        }
    }
    return RAPT_OK;
}

int rapt_b200_unshard_dev(int world, int period, const int32_t *offsets, int64_t n_max, int ncol, int64_t n_total,
                          const double *gathered, double *out, void *stream)
{
    if (int rc = ensure_init()) return rc;
    if (world < 1 || world > 64 || n_max < 0 || ncol < 1 || n_total < 0 || n_total > (int64_t)world * n_max || !gathered || !out)
        return fail(RAPT_E_ARG, "unshard: bad argument (1 <= world <= 64)");
    ShardTable tb;
    memset(&tb, 0, sizeof tb);
    tb.world = world;
    if (offsets) {
        if (period < world || offsets[0] != 0 || offsets[world] != period) return fail(RAPT_E_ARG, "unshard: offsets must run from 0 to period");
        for (int r = 0; r <= world; r++) {
            if (r && offsets[r] < offsets[r - 1]) return fail(RAPT_E_ARG, "unshard: offsets must ascend");
            tb.off[r] = offsets[r];
        }
        tb.period = period;
    } else {                                         // round-robin
        tb.period = world;
        for (int r = 0; r <= world; r++) tb.off[r] = r;
    }
    tb.full = n_total / tb.period;
    const long long rem = n_total - tb.full * tb.period;
    for (int r = 0; r <= world; r++)                 // ShardPlan._tail: proportional cut; round-robin: the first rem ranks
        tb.off2[r] = offsets ? (int)(((long long)tb.off[r] * rem) / tb.period) : (int)std::min<long long>(r, rem);
    if (n_total == 0) return RAPT_OK;
    if (int rc = check_on_bound_device(gathered, "unshard_dev")) return rc;
    const long long total = n_total * ncol;
    const int grid = (int)std::min<long long>((long long)g_sms * 16, (total + 255) / 256);
    k_unshard<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(tb, n_max, ncol, n_total, gathered, out);
    CK(cudaGetLastError());
    g_launches++;
    return RAPT_OK;
}

int rapt_b200_adaptive_last_stats(double *out, int n)
{
    if (!out || n < 15) return fail(RAPT_E_ARG, "adaptive_last_stats: need room for 15 doubles");
    const AdaptiveStats &a = g_adaptive_stats;
    const double v[15] = {(double)a.epochs, (double)a.launches_p, (double)a.launches_g, (double)a.tracer_launches_p,
                          (double)a.tracer_launches_g, (double)a.steps_p, (double)a.accepted_p, (double)a.calls_p,
                          (double)a.steps_g, (double)a.calls_g, a.ms_particle, a.ms_gc, a.ms_switch, a.ms_epochs, 0.0};
    for (int k = 0; k < 15; k++) out[k] = v[k];
    for (size_t k = 0; k < a.per_epoch.size() && 15 + (int)k < n; k++) out[15 + k] = a.per_epoch[k];
    return RAPT_OK;
}

int rapt_b200_final_diagnostics_dev(int kind, int64_t n, int ncol, const double *const *cols, const double *mass,
                                    const int32_t *status, double *packed, int nbins, double lo, double hi,
                                    int64_t *hist, double *stats, void *stream)
{
    if (int rc = ensure_init()) return rc;
    if (n < 0 || ncol < 4 || ncol > 8 || !cols || nbins < 1 || nbins > 8192 || !(hi > lo) || !hist || !stats)
        return fail(RAPT_E_ARG, "final_diagnostics: need 4 <= ncol <= 8 state columns, 1 <= nbins <= 8192, hi > lo, hist and stats");
    if (kind != 0 && kind != 1) return fail(RAPT_E_ARG, "final_diagnostics: kind %d (0 = log10 KE[eV], 1 = r / Re)", kind);
    if (kind == 0 && (ncol < 7 || !mass)) return fail(RAPT_E_ARG, "final_diagnostics: the kinetic-energy histogram needs 7 columns and mass");
    if (n == 0) return RAPT_OK;
    rapt::DiagCols c;
    for (int k = 0; k < 8; k++) c.col[k] = k < ncol ? cols[k] : nullptr;
    for (int k = 0; k < ncol; k++) if (!c.col[k]) return fail(RAPT_E_ARG, "final_diagnostics: null column %d", k);
    if (int rc = check_on_bound_device(c.col[0], "final_diagnostics_dev")) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = (int)std::min<long long>((long long)g_sms * 8, (n + 255) / 256);
    k_final_diagnostics<<<grid, 256, nbins * sizeof(unsigned int), s>>>(kind, n, ncol, c, mass, status, packed, nbins, lo, hi,
                                                                       reinterpret_cast<unsigned long long *>(hist), stats);
    CK(cudaGetLastError());
    g_launches++;
    return RAPT_OK;
}

}  // extern "C"
