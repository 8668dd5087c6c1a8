// evo_runtime.cu -- host runtime + C-ABI of the B200 multigrid evaluation library.
//
// Replaces, for one individual, the reference's  "emit ExaSlang -> java (x2) -> make -> run binary x
// samples -> parse stdout"  (evostencils/code_generation/exastencils.py:485-537) by
//   evo_cycle_build : op list -> bound kernel launches, captured once as a CUDA graph whose root is a
//                     WHILE conditional node (the generated solver's outer loop runs on the device)
//   evo_cycle_solve : one graph launch per sample, CUDA-event timed, residual history read back
// No CPU fallback exists: without a CUDA device evo_problem_create fails with EVO_ERR_NO_DEVICE.
#include "evo_runtime_internal.cuh"

// Many individuals are evaluated concurrently, one stream each: the default of 8 hardware work queues
// would serialise unrelated streams, so ask for the maximum before the CUDA context is created.
namespace {
struct EnvInit {
    EnvInit() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
} g_env_init;
}  // namespace

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// every live problem, so that an allocation failure can reclaim the recycled work slabs of all of them
static std::vector<evo_problem *> g_problems;

static Geom make_geom(int level, int dim)
{
    Geom g;
    g.n = (1 << level) + 1;
    g.dim = dim;
    g.pitch = (g.n + 15) / 16 * 16;
    g.nz = dim == 3 ? g.n : 1;
    g.plane = (long long)g.pitch * g.n;
    g.total = g.plane * g.nz;
    g.zlo = g.zin0 = 1;
    g.zhi = g.zin1 = g.n - 2;
    g.zpar = 0;
    g.zoff = 0;
    return g;
}

static TransferW make_transfer(const double *w, int dim)
{
    TransferW t;
    memset(&t, 0, sizeof(t));
    for (int p = 0; p < 27; ++p) {
        if (w[p] == 0.0) continue;
        int ox = p % 3 - 1, oy = (p / 3) % 3 - 1, oz = p / 9 - 1;
        if (dim == 2 && oz != 0) continue;
        int q = t.nnz++;
        t.ox[q] = (signed char)ox; t.oy[q] = (signed char)oy; t.oz[q] = (signed char)oz;
        t.w[q] = w[p];
    }
    return t;
}

// ------------------------------------------------------------------------------------------------
// library
extern "C" int evo_abi_version(void) { return EVO_ABI_VERSION; }
extern "C" const char *evo_last_error(void) { return g_err.c_str(); }

// tuning switches: `name` is the environment-variable name (EVO_RB_VARIANT, EVO_RB_FUSE2, ...)
extern "C" int evo_set_option(const char *name, int value)
{
    if (!name) return fail(EVO_ERR_INVALID, "null argument");
    for (int id = 0; id < OPT_COUNT; ++id)
        if (strcmp(name, option_name(id)) == 0) {
            option_table().value[id] = value;
            option_table().init[id] = true;
            return EVO_OK;
        }
    return fail(EVO_ERR_INVALID, "unknown option %s", name);
}

extern "C" int evo_get_option(const char *name, int *value)
{
    if (!name || !value) return fail(EVO_ERR_INVALID, "null argument");
    for (int id = 0; id < OPT_COUNT; ++id)
        if (strcmp(name, option_name(id)) == 0) { *value = option(id); return EVO_OK; }
    return fail(EVO_ERR_INVALID, "unknown option %s", name);
}

extern "C" int evo_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        fail(EVO_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return EVO_ERR_NO_DEVICE;
    }
    return n;
}

extern "C" int evo_device_name(int device, char *buf, size_t len, int *sm_count)
{
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (buf && len) snprintf(buf, len, "%s", prop.name);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return EVO_OK;
}

// ------------------------------------------------------------------------------------------------
// problem
extern "C" int evo_problem_create(const evo_problem_desc *d, evo_problem **out)
{
    if (!d || !out) return fail(EVO_ERR_INVALID, "null argument");
    if (d->abi_version != EVO_ABI_VERSION) return fail(EVO_ERR_INVALID, "ABI version mismatch");
    if (d->dim < 2 || d->dim > 3 || d->n_fields < 1 || d->n_fields > EVO_MAX_FIELDS || d->min_level < 1 ||
        d->max_level >= EVO_MAX_LEVELS || d->min_level > d->max_level || (d->scalar_words != 1 && d->scalar_words != 2))
        return fail(EVO_ERR_INVALID, "invalid problem descriptor");
    if (d->scalar_words == 2 && (d->n_fields != 1 || d->dim != 2))
        return fail(EVO_ERR_UNSUPPORTED, "complex problems: scalar 2-D only");
    int ndev = evo_device_count();
    if (ndev <= 0) return fail(EVO_ERR_NO_DEVICE, "no CUDA device: this backend has no CPU fallback");
    if (d->device < 0 || d->device >= ndev) return fail(EVO_ERR_INVALID, "device ordinal %d out of range", d->device);
    CU(cudaSetDevice(d->device));
    evo_problem *p = new evo_problem();
    p->desc = *d;
    p->words = d->scalar_words;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, d->device));
    p->sm_count = prop.multiProcessorCount;
    for (int l = d->min_level; l <= d->max_level; ++l) p->geom[l] = make_geom(l, d->dim);
    p->R = make_transfer(d->restrict_w, d->dim);
    p->P = make_transfer(d->prolong_w, d->dim);
    const size_t bytes = (size_t)p->geom[d->max_level].total * sizeof(double) * p->words;
    for (int i = 0; i < EVO_MAX_FIELDS; ++i) { p->init_sol[i] = nullptr; p->rhs0[i] = nullptr; }
    for (int i = 0; i < d->n_fields; ++i) {
        CU(cudaMalloc(&p->init_sol[i], bytes));
        CU(cudaMalloc(&p->rhs0[i], bytes));
        CU(cudaMemset(p->init_sol[i], 0, bytes));
        CU(cudaMemset(p->rhs0[i], 0, bytes));
    }
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { total_b = (size_t)64 << 30; cudaGetLastError(); }
    p->device_bytes = total_b;
    g_problems.push_back(p);
    *out = p;
    return EVO_OK;
}

static void free_problem(evo_problem *p);

extern "C" int evo_problem_destroy(evo_problem *p)
{
    if (!p) return EVO_OK;
    if (p->live_cycles > 0) { p->closed = true; return EVO_OK; }   // freed when its last cycle goes
    free_problem(p);
    return EVO_OK;
}

static void free_problem(evo_problem *p)
{
    g_problems.erase(std::remove(g_problems.begin(), g_problems.end(), p), g_problems.end());
    cudaSetDevice(p->desc.device);
    for (int i = 0; i < EVO_MAX_FIELDS; ++i) { cudaFree(p->init_sol[i]); cudaFree(p->rhs0[i]); }
    for (auto &s : p->pool) cudaFree(s.first);
    for (auto &cr : p->res_pool) {
        if (cr.d_hist) cudaFree(cr.d_hist);
        if (cr.h_state) cudaFreeHost(cr.h_state);
        if (cr.h_hist) cudaFreeHost(cr.h_hist);
        cudaEventDestroy(cr.ev0); cudaEventDestroy(cr.ev1); cudaStreamDestroy(cr.stream);
    }
    delete p;
}

extern "C" int evo_problem_set_slab_ex(evo_problem *p, int rank, int world, int lc, int ghost)
{
    if (!p) return fail(EVO_ERR_INVALID, "null argument");
    const evo_problem_desc &d = p->desc;
    if (d.dim != 3 || d.n_fields != 1 || d.scalar_words != 1 || d.kind != EVO_PROBLEM_LINEAR)
        return fail(EVO_ERR_UNSUPPORTED, "domain decomposition: 3-D real scalar linear problems only");
    if (world < 1 || rank < 0 || rank >= world) return fail(EVO_ERR_INVALID, "invalid rank/world");
    if (ghost < 2 || ghost > 16 || (ghost & 1)) return fail(EVO_ERR_INVALID, "ghost planes per side: an even number in [2, 16]");
    if (lc < 5 || lc < d.min_level || lc > d.max_level) return fail(EVO_ERR_INVALID, "coarsest distributed level must be >= 5 and within the hierarchy");
    const int inner = (1 << lc) - 1;
    if (inner < ghost * world) return fail(EVO_ERR_INVALID, "too many ranks for the coarsest distributed level");
    if (p->live_cycles > 0) return fail(EVO_ERR_INVALID, "set the slab before building cycles");
    CU(cudaSetDevice(d.device));
    p->slab_world = world; p->slab_rank = rank; p->slab_lc = lc;
    const int base = inner / world, rem = inner % world;
    int a = 1 + rank * base + std::min(rank, rem), b = a + base + (rank < rem ? 1 : 0) - 1;
    const int G = ghost;
    for (int l = lc; l <= d.max_level; ++l) {
        if (l > lc) { a = 2 * a - 1; b = 2 * b + (rank == world - 1 ? 1 : 0); }
        Geom g = make_geom(l, 3);
        p->own_g0[l] = a; p->own_g1[l] = b;
        g.zoff = a - G;
        g.nz = (b - a + 1) + 2 * G;
        g.zlo = G; g.zhi = G + (b - a);
        g.zin0 = std::max(0, 1 - g.zoff);
        g.zin1 = std::min(g.nz - 1, (g.n - 2) - g.zoff);
        g.zpar = g.zoff & 1;
        g.total = g.plane * g.nz;
        p->geom[l] = g;
    }
    // re-allocate the finest-level initial fields for the local slab
    const size_t bytes = (size_t)p->geom[d.max_level].total * sizeof(double);
    for (int i = 0; i < d.n_fields; ++i) {
        cudaFree(p->init_sol[i]); cudaFree(p->rhs0[i]);
        CU(cudaMalloc(&p->init_sol[i], bytes));
        CU(cudaMalloc(&p->rhs0[i], bytes));
        CU(cudaMemset(p->init_sol[i], 0, bytes));
        CU(cudaMemset(p->rhs0[i], 0, bytes));
    }
    return EVO_OK;
}

extern "C" int evo_problem_set_slab(evo_problem *p, int rank, int world, int lc) { return evo_problem_set_slab_ex(p, rank, world, lc, 2); }

extern "C" int evo_problem_slab_info(evo_problem *p, int level, long long info[8])
{
    if (!p || !info || level < p->desc.min_level || level > p->desc.max_level) return fail(EVO_ERR_INVALID, "invalid argument");
    const Geom &g = p->geom[level];
    info[0] = g.zoff; info[1] = g.nz; info[2] = g.zlo; info[3] = g.zhi;
    info[4] = g.zlo + g.zoff; info[5] = g.zhi + g.zoff; info[6] = g.pitch; info[7] = g.plane;
    return EVO_OK;
}

// dense host array (n^dim entries, x fastest) <-> padded device array
static int copy_field(const Geom &g, int words, void *dev, const void *host_src, void *host_dst, cudaStream_t stream)
{
    const size_t esz = sizeof(double) * words;
    // local planes that exist in the global grid (a z-slab holds planes [zoff, zoff + nz))
    const int gn = g.dim == 3 ? g.n : 1;
    int l0 = 0, l1 = g.nz - 1;                       // upload: every local plane incl. ghosts
    if (host_dst && g.dim == 3 && (g.zoff != 0 || g.nz != g.n)) { l0 = g.zlo; l1 = g.zhi; }   // download: owned planes
    l0 = std::max(l0, -g.zoff);
    l1 = std::min(l1, gn - 1 - g.zoff);
    if (l1 < l0) return EVO_OK;
    const size_t rows = (size_t)g.n * (l1 - l0 + 1);
    char *d0 = (char *)dev + (size_t)l0 * g.plane * esz;
    const size_t hoff = (size_t)(l0 + g.zoff) * g.n * g.n * esz;
    if (host_src) CU(cudaMemcpy2DAsync(d0, g.pitch * esz, (const char *)host_src + hoff, g.n * esz, g.n * esz, rows, cudaMemcpyHostToDevice, stream));
    if (host_dst) CU(cudaMemcpy2DAsync((char *)host_dst + hoff, g.n * esz, d0, g.pitch * esz, g.n * esz, rows, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    return EVO_OK;
}

extern "C" int evo_problem_set_field(evo_problem *p, int level, int buf, int field, const double *host, size_t n_doubles)
{
    if (!p || !host) return fail(EVO_ERR_INVALID, "null argument");
    if (level != p->desc.max_level) return fail(EVO_ERR_UNSUPPORTED, "initial fields are defined on the finest level only");
    if (field < 0 || field >= p->desc.n_fields || (buf != EVO_BUF_SOL && buf != EVO_BUF_RHS))
        return fail(EVO_ERR_INVALID, "invalid field/buffer");
    const Geom &g = p->geom[level];
    if (n_doubles != (size_t)g.n * g.n * (g.dim == 3 ? g.n : 1) * p->words) return fail(EVO_ERR_INVALID, "field size mismatch");
    CU(cudaSetDevice(p->desc.device));
    return copy_field(g, p->words, buf == EVO_BUF_SOL ? p->init_sol[field] : p->rhs0[field], host, nullptr, 0);
}

// ------------------------------------------------------------------------------------------------
// cycle: memory
// pointwise RB-GS on a 3-D scalar real problem runs as the out-of-place streaming kernel
static bool rb_stream_candidate(const evo_cycle *c, const evo_op &op)
{
    const evo_problem_desc &d = c->p->desc;
    // (3-D: k3_rbgs_col from 33^3; 2-D: the register-streamed warp kernels from 129^2)
    return op.code == EVO_OP_SMOOTH && op.mode == EVO_SMOOTH_REDBLACK && op.kind == EVO_KIND_LINEAR && op.n_unknowns == 1 &&
           d.n_fields == 1 && d.scalar_words == 1 && d.kind == EVO_PROBLEM_LINEAR &&
           ((d.dim == 3 && c->p->geom[op.level].n >= 33) || (d.dim == 2 && c->p->geom[op.level].n >= w2::W2_MIN_N));
}
static bool level_needs_slot(const evo_cycle *c, int level)
{
    for (const evo_op &op : c->ops)
        if (op.level == level && ((op.code == EVO_OP_SMOOTH && op.mode == EVO_SMOOTH_JACOBI) || op.code == EVO_OP_RICHARDSON ||
                                  rb_stream_candidate(c, op) ||
                                  (op.code == EVO_OP_COARSE_SOLVE && c->p->desc.kind == EVO_PROBLEM_FAS)))
            return true;
    return false;
}
static bool uses_buffer(const evo_cycle *c, int level, int buf)
{
    for (const evo_op &op : c->ops) {
        int dl = op.level, sl = op.level;
        if (op.code == EVO_OP_RESTRICT) dl = op.level - 1;
        if (op.code == EVO_OP_PROLONG_ADD || op.code == EVO_OP_PROLONG_SET) sl = op.level - 1;
        if ((dl == level && op.dst == buf) || (sl == level && op.src == buf)) return true;
    }
    return false;
}

static void release_pools(int device)
{
    for (evo_problem *q : g_problems) {
        if (q->desc.device != device) continue;
        for (auto &sl : q->pool) cudaFree(sl.first);
        q->pool.clear();
    }
}

static int allocate_cycle(evo_cycle *c)
{
    evo_problem *p = c->p;
    const int nf = p->desc.n_fields, lo = p->desc.min_level, hi = p->desc.max_level;
    const size_t esz = sizeof(double) * p->words;
    auto align = [](size_t b) { return (b + 255) / 256 * 256; };
    // pass 1: size, pass 2: carve
    for (int pass = 0; pass < 2; ++pass) {
        size_t off = 0;
        char *base = (char *)c->slab;
        auto take = [&](size_t bytes) -> void * {
            void *r = pass ? (void *)(base + off) : nullptr;
            off += align(bytes);
            return r;
        };
        for (int l = lo; l <= hi; ++l) {
            const size_t fb = (size_t)p->geom[l].total * esz;
            const bool slot = level_needs_slot(c, l);
            for (int i = 0; i < nf; ++i) {
                c->lv[l].buf[EVO_BUF_SOL][i] = take(fb);
                c->lv[l].buf[EVO_BUF_RHS][i] = (l == hi && p->desc.kind != EVO_PROBLEM_HELMHOLTZ)
                                                   ? (pass ? p->rhs0[i] : nullptr) : take(fb);   // Helmholtz: f is rewritten per application
                c->lv[l].buf[EVO_BUF_RES][i] = take(fb);
                c->lv[l].buf[EVO_BUF_COR][i] = (l == hi && uses_buffer(c, l, EVO_BUF_COR)) ? take(fb) : c->lv[l].buf[EVO_BUF_SOL][i];
                c->lv[l].buf[EVO_BUF_APX][i] = (p->desc.kind == EVO_PROBLEM_FAS) ? take(fb) : nullptr;
                c->lv[l].slot[i] = slot ? take(fb) : nullptr;
                c->lv[l].swapped[i] = false;
            }
        }
        const size_t cb = (size_t)p->geom[lo].total * esz;
        for (int v = 0; v < 8; ++v)
            for (int i = 0; i < nf; ++i) c->krylov[v][i] = take(cb);
        if (p->desc.kind == EVO_PROBLEM_HELMHOLTZ) {
            const size_t fbh = (size_t)p->geom[hi].total * esz;
            for (int v = 0; v < 9; ++v) c->helm[v] = take(fbh);
            c->d_helm = (helm::HelmState *)take(sizeof(helm::HelmState));
        }
        c->d_state = (SolveState *)take(sizeof(SolveState));
        c->d_cg_iters = (int *)take(256);
        const Geom &gf = p->geom[hi];
        c->n_partials = nf * (gf.n - 2) * (gf.dim == 3 ? gf.n - 2 : 1);
        c->d_partials = (double *)take(sizeof(double) * ((size_t)c->n_partials + (size_t)nf * gf.n));
        // constant tables of the fused runs (evo_kernels_run.cuh): behind everything a reset clears
        const size_t zero_end = off;
        c->d_run_sten = (OpSten *)take(sizeof(OpSten) * EVO_MAX_LEVELS);
        c->d_run_sp = (SmoothParams *)take(sizeof(SmoothParams) * std::max<size_t>(1, c->ops.size()));
        c->d_run_rp = (TransferW *)take(sizeof(TransferW) * 2);
        if (pass == 0) {
            c->zero_bytes = zero_end;
            c->slab_bytes = off;
            c->slab_cap = off;
            c->slab = nullptr;
            // best fit among the recycled slabs: the smallest one that is large enough and wastes at most 25 %
            size_t best = (size_t)-1;
            for (size_t k = 0; k < p->pool.size(); ++k)
                if (p->pool[k].second >= off && p->pool[k].second <= off + off / 4 &&
                    (best == (size_t)-1 || p->pool[k].second < p->pool[best].second))
                    best = k;
            if (best != (size_t)-1) {
                c->slab = p->pool[best].first;
                c->slab_cap = p->pool[best].second;
                p->pool.erase(p->pool.begin() + best);
            }
            if (!c->slab) {
                cudaError_t e = cudaMalloc(&c->slab, off);
                if (e == cudaErrorMemoryAllocation) {
                    // idle slabs of other sizes may hold the memory: give every pool on this device back and retry
                    cudaGetLastError();
                    c->slab = nullptr;
                    release_pools(p->desc.device);
                    e = cudaMalloc(&c->slab, off);
                }
                if (e != cudaSuccess) {
                    c->slab = nullptr;
                    cudaGetLastError();
                    return fail(e == cudaErrorMemoryAllocation ? EVO_ERR_OOM : EVO_ERR_CUDA, "cudaMalloc of a %zu-byte work slab: %s", off,
                                cudaGetErrorString(e));
                }
            }
        }
    }
    return EVO_OK;
}

static int reset_cycle(evo_cycle *c, cudaStream_t s)
{
    evo_problem *p = c->p;
    const int nf = p->desc.n_fields, hi = p->desc.max_level;
    const size_t esz = sizeof(double) * p->words;
    // zero everything we own (boundary layers of the coarse fields must be 0 and are never written) -- except the
    // finest SOL arrays and their [next] slots, which are overwritten by the initial guess right below
    const size_t fb = (size_t)p->geom[hi].total * esz;
    std::vector<std::pair<char *, size_t>> skip;
    for (int i = 0; i < nf; ++i) {
        skip.push_back({(char *)c->lv[hi].buf[EVO_BUF_SOL][i], fb});
        if (c->lv[hi].slot[i]) skip.push_back({(char *)c->lv[hi].slot[i], fb});
    }
    std::sort(skip.begin(), skip.end());
    char *cur = (char *)c->slab, *end = (char *)c->slab + c->zero_bytes;
    for (const auto &sk : skip) {
        if (sk.first > cur) CU(cudaMemsetAsync(cur, 0, (size_t)(sk.first - cur), s));
        cur = std::max(cur, sk.first + sk.second);
    }
    if (end > cur) CU(cudaMemsetAsync(cur, 0, (size_t)(end - cur), s));
    for (int i = 0; i < nf; ++i) {
        CU(cudaMemcpyAsync(c->lv[hi].buf[EVO_BUF_SOL][i], p->init_sol[i], fb, cudaMemcpyDeviceToDevice, s));
        // both jacobi slots carry the Dirichlet boundary values
        if (c->lv[hi].slot[i]) CU(cudaMemcpyAsync(c->lv[hi].slot[i], p->init_sol[i], fb, cudaMemcpyDeviceToDevice, s));
    }
    c->pristine = true;
    return EVO_OK;
}

// the statement dispatch lives in evo_dispatch_inst.cu (one translation unit per combination)
#define EVO_EXTERN_INST(...)                                                                  \
    extern template int enqueue_op<__VA_ARGS__>(evo_cycle *, const evo_op &, cudaStream_t);    \
    extern template int op_residual<__VA_ARGS__>(evo_cycle *, int, bool, cudaStream_t);        \
    extern template int op_restrict<__VA_ARGS__>(evo_cycle *, const evo_op &, cudaStream_t);   \
    extern template int op_reduce_rows<__VA_ARGS__>(evo_cycle *, int, cudaStream_t);     \
    extern template bool run_eligible<__VA_ARGS__>(const evo_cycle *, const evo_op &);      \
    extern template int enqueue_run<__VA_ARGS__>(evo_cycle *, const evo_op *, int, cudaStream_t);
EVO_EXTERN_INST(double, 2, 1)
EVO_EXTERN_INST(double, 2, 2)
EVO_EXTERN_INST(double, 3, 1)
EVO_EXTERN_INST(double, 3, 2)
EVO_EXTERN_INST(cplx, 2, 1)

// ---- FAS problems (real scalar 2-D) -----------------------------------------------------------------
static int fas_residual(evo_cycle *c, int l, cudaStream_t s)
{
    fas::Lin2 L;
    if (!fas::make_lin2(c->sten[l].s[0][0], &L)) return fail(EVO_ERR_UNSUPPORTED, "FAS: unsupported linear stencil");
    const Geom &g = c->p->geom[l];
    fas::k2_fas_residual<<<row_grid(g), BX, 0, s>>>(g, L, c->p->desc.gamma, (const double *)c->lv[l].buf[EVO_BUF_SOL][0],
                                                   (const double *)c->lv[l].buf[EVO_BUF_RHS][0], (double *)c->lv[l].buf[EVO_BUF_RES][0]);
    c->launch_counter++;
    CU(cudaGetLastError());
    return EVO_OK;
}

// returns 1 when the op is not FAS specific
static int fas_dispatch(evo_cycle *c, const evo_op &op, cudaStream_t s)
{
    const evo_problem_desc &d = c->p->desc;
    const int l = op.level;
    const double gamma = d.gamma;
    auto swap_slot = [&](int lev) {
        bool cor_alias = c->lv[lev].buf[EVO_BUF_COR][0] == c->lv[lev].buf[EVO_BUF_SOL][0];
        std::swap(c->lv[lev].buf[EVO_BUF_SOL][0], c->lv[lev].slot[0]);
        if (cor_alias) c->lv[lev].buf[EVO_BUF_COR][0] = c->lv[lev].buf[EVO_BUF_SOL][0];
        c->lv[lev].swapped[0] = !c->lv[lev].swapped[0];
    };
    switch (op.code) {
    case EVO_OP_SMOOTH: {
        if (op.kind == EVO_KIND_LINEAR) return 1;
        fas::Lin2 L;
        if (!fas::make_lin2(c->sten[l].s[0][0], &L)) return fail(EVO_ERR_UNSUPPORTED, "FAS: unsupported linear stencil");
        const Geom &g = c->p->geom[l];
        const int newton = op.kind == EVO_KIND_FAS_NEWTON, steps = op.count > 0 ? op.count : 1;
        const double *f = (const double *)c->lv[l].buf[EVO_BUF_RHS][0];
        if (op.mode == EVO_SMOOTH_JACOBI) {
            if (!c->lv[l].slot[0]) return fail(EVO_ERR_INVALID, "missing jacobi slot");
            w2::Star5 c5;
            if (w2::star5_applicable(g, c->sten[l]) && w2::match_star5(c->sten[l].s[0][0], &c5)) {
                // large grids: register-streamed warp kernel (evo_kernels_warp2d.cuh)
                w2::FasPoint upd{c5, gamma, op.omega, newton, steps};
                if (!w2::launch_sweep<w2::FasPoint, 1, false>(c->p->sm_count, g, upd, (const double *)c->lv[l].buf[EVO_BUF_SOL][0], f,
                                                              (double *)c->lv[l].slot[0], s))
                    return fail(EVO_ERR_CUDA, "FAS streaming sweep: launch failed");
                c->launch_counter++;
                swap_slot(l);
                return EVO_OK;
            }
            fas::k2_fas_smooth<<<row_grid(g), BX, 0, s>>>(g, L, gamma, (const double *)c->lv[l].buf[EVO_BUF_SOL][0],
                                                         (double *)c->lv[l].slot[0], f, newton, steps, op.omega, -1);
            c->launch_counter++;
            swap_slot(l);
        } else if (op.mode == EVO_SMOOTH_REDBLACK) {
            double *u = (double *)c->lv[l].buf[EVO_BUF_SOL][0];
            for (int color = 0; color < 2; ++color) {
                fas::k2_fas_smooth<<<row_grid(g, 2), BX, 0, s>>>(g, L, gamma, u, u, f, newton, steps, op.omega, color);
                c->launch_counter++;
            }
        } else {
            return fail(EVO_ERR_UNSUPPORTED, "FAS: lexicographic smoothing is not implemented");
        }
        CU(cudaGetLastError());
        return EVO_OK;
    }
    case EVO_OP_RESIDUAL: return fas_residual(c, l, s);
    case EVO_OP_FAS_RESTRICT_SOL: {
        evo_op r = op;
        r.code = EVO_OP_RESTRICT; r.dst = EVO_BUF_APX; r.src = EVO_BUF_SOL;
        EV((op_restrict<double, 2, 1>(c, r, s)));
        const size_t bytes = (size_t)c->p->geom[l - 1].total * sizeof(double);
        CU(cudaMemcpyAsync(c->lv[l - 1].buf[EVO_BUF_SOL][0], c->lv[l - 1].buf[EVO_BUF_APX][0], bytes, cudaMemcpyDeviceToDevice, s));
        return EVO_OK;
    }
    case EVO_OP_FAS_COARSE_RHS: {
        evo_op r = op;
        r.code = EVO_OP_RESTRICT; r.dst = EVO_BUF_RHS; r.src = EVO_BUF_RES;
        EV((op_restrict<double, 2, 1>(c, r, s)));
        fas::Lin2 L;
        if (!c->has_sten[l - 1] || !fas::make_lin2(c->sten[l - 1].s[0][0], &L)) return fail(EVO_ERR_INVALID, "FAS: no operator on level %d", l - 1);
        const Geom &gc = c->p->geom[l - 1];
        fas::k2_fas_add_operator<<<row_grid(gc), BX, 0, s>>>(gc, L, gamma, (const double *)c->lv[l - 1].buf[EVO_BUF_APX][0],
                                                            (double *)c->lv[l - 1].buf[EVO_BUF_RHS][0]);
        c->launch_counter++;
        CU(cudaGetLastError());
        return EVO_OK;
    }
    case EVO_OP_FAS_SUB_APX: {
        const long long n = c->p->geom[l].total;
        fas::k_sub_inplace<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, s>>>(
            (double *)c->lv[l].buf[EVO_BUF_SOL][0], (const double *)c->lv[l].buf[EVO_BUF_APX][0], n);
        c->launch_counter++;
        CU(cudaGetLastError());
        return EVO_OK;
    }
    case EVO_OP_COARSE_SOLVE: {
        fas::Lin2 L;
        if (!fas::make_lin2(c->sten[l].s[0][0], &L)) return fail(EVO_ERR_UNSUPPORTED, "FAS: unsupported linear stencil");
        if (!c->lv[l].slot[0]) return fail(EVO_ERR_INVALID, "missing slot for the FAS coarse solver");
        const Geom &g = c->p->geom[l];
        const int cgs_mode = option(OPT_FAS_CGS);   // 0 auto, 1 one CTA (smem), 2 launches
        if (cgs_mode == 0 && g.n <= 65) {   // 129^2: 200 graph-launched sweeps over all SMs are faster (measured)
            // rows split over a cluster of 8 CTAs (distributed shared memory halos, one cluster barrier per sweep)
            const int cs = 8, ni = g.n - 2, rp = (ni + cs - 1) / cs;
            const int npt = (rp * ni + 511) / 512;
            const size_t smem = (size_t)2 * (rp + 2) * (g.n + 1) * sizeof(double);
            auto kern = npt <= 1 ? fas::k2_fas_coarse_cluster<1> : fas::k2_fas_coarse_cluster<4>;
            if (npt <= 4) {
                static bool attr1 = false, attr4 = false;
                bool &attr = npt <= 1 ? attr1 : attr4;
                if (!attr) {
                    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
                    attr = true;
                }
                cudaLaunchConfig_t cfg;
                memset(&cfg, 0, sizeof(cfg));
                cfg.gridDim = dim3(cs);
                cfg.blockDim = dim3(512);
                cfg.dynamicSmemBytes = smem;
                cfg.stream = s;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                double *ua = (double *)c->lv[l].buf[EVO_BUF_SOL][0];
                const double *fa = (const double *)c->lv[l].buf[EVO_BUF_RHS][0];
                int cnt = op.count;
                double om = op.omega, gm = gamma;
                CU(cudaLaunchKernelEx(&cfg, kern, g, L, gm, ua, fa, cnt, om));
                c->launch_counter++;
                return EVO_OK;
            }
        }
        if ((long long)(g.n - 2) * (g.n - 2) > 4096 || cgs_mode == 2) {
            // a "coarsest" grid too large for one CTA (BASELINE config 4097^2 has 257^2 there): one launch per sweep
            // over all SMs, ping-pong between the two SOL slots; the per-node arithmetic is that of k2_fas_coarse
            const double *f = (const double *)c->lv[l].buf[EVO_BUF_RHS][0];
            w2::Star5 c5;
            // (a coarsest grid below ~1025^2 is latency bound: one launch per sweep over all SMs is faster than fused sweeps)
            const bool stream = g.n >= 1025 && w2::star5_applicable(g, c->sten[l]) && w2::match_star5(c->sten[l].s[0][0], &c5);
            for (int t = 0; t < op.count;) {
                if (stream) {
                    // four (two, one) Newton-Jacobi sweeps per launch: the pipeline stages of the streaming kernel
                    w2::FasPoint upd{c5, gamma, op.omega, 1, 1};
                    const int k = op.count - t >= 4 ? 4 : (op.count - t >= 2 ? 2 : 1);
                    const double *src = (const double *)c->lv[l].buf[EVO_BUF_SOL][0];
                    double *dst = (double *)c->lv[l].slot[0];
                    bool ok;
                    if (k == 4) ok = w2::launch_sweep<w2::FasPoint, 4, false>(c->p->sm_count, g, upd, src, f, dst, s);
                    else if (k == 2) ok = w2::launch_sweep<w2::FasPoint, 2, false>(c->p->sm_count, g, upd, src, f, dst, s);
                    else ok = w2::launch_sweep<w2::FasPoint, 1, false>(c->p->sm_count, g, upd, src, f, dst, s);
                    if (!ok) return fail(EVO_ERR_CUDA, "FAS streaming sweep: launch failed");
                    t += k;
                } else {
                    fas::k2_fas_smooth<<<row_grid(g), BX, 0, s>>>(g, L, gamma, (const double *)c->lv[l].buf[EVO_BUF_SOL][0],
                                                                 (double *)c->lv[l].slot[0], f, 1, 1, op.omega, -1);
                    ++t;
                }
                c->launch_counter++;
                swap_slot(l);
            }
            CU(cudaGetLastError());
            return EVO_OK;
        }
        const bool global_cgs = option(OPT_FAS_CGS_GLOBAL) != 0;
        if (!global_cgs && g.n <= 65) {
            // both slots in shared memory, up to 4 nodes per thread in flight
            const size_t smem = (size_t)2 * (g.n + 1) * g.n * sizeof(double);
            static bool attr = false;
            if (!attr) {
                CU(cudaFuncSetAttribute(fas::k2_fas_coarse_smem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 66 * 65 * 8));
                attr = true;
            }
            fas::k2_fas_coarse_smem<4><<<1, 1024, smem, s>>>(g, L, gamma, (double *)c->lv[l].buf[EVO_BUF_SOL][0],
                                                             (const double *)c->lv[l].buf[EVO_BUF_RHS][0], op.count, op.omega);
            c->launch_counter++;
            CU(cudaGetLastError());
            return EVO_OK;
        }
        fas::k2_fas_coarse<<<1, 1024, 0, s>>>(g, L, gamma, (double *)c->lv[l].buf[EVO_BUF_SOL][0],
                                              (double *)c->lv[l].slot[0], (const double *)c->lv[l].buf[EVO_BUF_RHS][0], op.count, op.omega);
        c->launch_counter++;
        CU(cudaGetLastError());
        return EVO_OK;
    }
    default: return 1;
    }
}

// 1 / (1 - i k h) by Smith's algorithm, the same operation sequence as the oracle's cdiv_smith(1.0, den)
cplx helm_rden(const evo_problem *p, int level)
{
    const double h = 1.0 / (double)(1 << level);
    // I * k = (0*kr - 1*ki) + (0*ki + 1*kr) i ; times h ; den = 1.0 - that
    const double kr = p->desc.k_re, ki = p->desc.k_im;
    const double ikr = 0.0 * kr - 1.0 * ki, iki = 0.0 * ki + 1.0 * kr;
    const cplx den(1.0 - ikr * h, 0.0 - iki * h);
    return cplx(1.0, 0.0) / den;
}

static int helm_bc(evo_cycle *c, int level, void *field, cudaStream_t s)
{
    const Geom &g = c->p->geom[level];
    helm::k2_helm_bc<<<(g.n + 127) / 128, 128, 0, s>>>(g, (cplx *)field, helm_rden(c->p, level));
    c->launch_counter++;
    CU(cudaGetLastError());
    return EVO_OK;
}

static int dispatch_op_inner(evo_cycle *c, const evo_op &op, cudaStream_t s);

// Helmholtz: u / gen_error_u / gen_residual_u carry the Robin boundary function; it is applied after every
// statement that writes one of them (same rule as oracle helm_bc_after_op)
static int dispatch_op(evo_cycle *c, const evo_op &op, cudaStream_t s)
{
    c->pristine = false;
    EV(dispatch_op_inner(c, op, s));
    if (c->p->desc.kind != EVO_PROBLEM_HELMHOLTZ) return EVO_OK;
    int buf = -1;
    switch (op.code) {
    case EVO_OP_ZERO: case EVO_OP_COPY: case EVO_OP_PROLONG_SET: buf = op.dst; break;
    case EVO_OP_SMOOTH: case EVO_OP_RICHARDSON: case EVO_OP_PROLONG_ADD: case EVO_OP_COARSE_SOLVE: buf = EVO_BUF_SOL; break;
    case EVO_OP_RESIDUAL: buf = EVO_BUF_RES; break;
    default: return EVO_OK;
    }
    if (buf == EVO_BUF_RHS) return EVO_OK;
    return helm_bc(c, op.level, c->lv[op.level].buf[buf][0], s);
}

static int dispatch_op_inner(evo_cycle *c, const evo_op &op, cudaStream_t s)
{
    const evo_problem_desc &d = c->p->desc;
    if (d.kind == EVO_PROBLEM_FAS) {
        if (d.dim != 2 || d.n_fields != 1 || d.scalar_words != 1) return fail(EVO_ERR_UNSUPPORTED, "FAS: real scalar 2-D only");
        int rc = fas_dispatch(c, op, s);
        if (rc != 1) return rc;
    }
    if (d.scalar_words == 2) return enqueue_op<cplx, 2, 1>(c, op, s);
    if (d.dim == 2) return d.n_fields == 1 ? enqueue_op<double, 2, 1>(c, op, s) : enqueue_op<double, 2, 2>(c, op, s);
    return d.n_fields == 1 ? enqueue_op<double, 3, 1>(c, op, s) : enqueue_op<double, 3, 2>(c, op, s);
}

static int dispatch_residual_norm(evo_cycle *c, cudaStream_t s, bool force_store = false)
{
    c->pristine = false;
    const bool saved_dead = c->res_dead_on_entry;
    if (force_store) c->res_dead_on_entry = false;
    struct Restore { evo_cycle *c; bool v; ~Restore() { c->res_dead_on_entry = v; } } restore{c, saved_dead};
    const evo_problem_desc &d = c->p->desc;
    const int l = d.max_level;
    if (d.kind == EVO_PROBLEM_FAS) {
        EV(fas_residual(c, l, s));
        const Geom &g = c->p->geom[l];
        const int ni = g.n - 2;
        auto r = fields_of<double>(c->lv[l].buf[EVO_BUF_RES], 1);
        k_row_sumsq<double, 2, 1><<<(unsigned)((ni + 3) / 4), 128, 0, s>>>(g, r, c->d_partials);
        c->launch_counter++;
        return op_reduce_rows<double, 2, 1>(c, ni, s);
    }
    if (d.scalar_words == 2) EV((op_residual<cplx, 2, 1>(c, l, true, s)));
    else if (d.dim == 2 && d.n_fields == 1) EV((op_residual<double, 2, 1>(c, l, true, s)));
    else if (d.dim == 2) EV((op_residual<double, 2, 2>(c, l, true, s)));
    else if (d.n_fields == 1) EV((op_residual<double, 3, 1>(c, l, true, s)));
    else EV((op_residual<double, 3, 2>(c, l, true, s)));
    return EVO_OK;
}

// all statements of the cycle function, then restore the canonical jacobi-slot assignment so that a
// replay (next outer iteration) starts from the same pointers
static bool op_run_eligible(const evo_cycle *c, const evo_op &op)
{
    const evo_problem_desc &d = c->p->desc;
    if (d.kind != EVO_PROBLEM_LINEAR || d.scalar_words != 1) return false;
    if (d.dim == 2) return d.n_fields == 1 ? run_eligible<double, 2, 1>(c, op) : run_eligible<double, 2, 2>(c, op);
    return d.n_fields == 1 ? run_eligible<double, 3, 1>(c, op) : run_eligible<double, 3, 2>(c, op);
}

static int enqueue_fused_run(evo_cycle *c, const evo_op *ops, int n, cudaStream_t s)
{
    const evo_problem_desc &d = c->p->desc;
    c->pristine = false;
    if (d.dim == 2) return d.n_fields == 1 ? enqueue_run<double, 2, 1>(c, ops, n, s) : enqueue_run<double, 2, 2>(c, ops, n, s);
    return d.n_fields == 1 ? enqueue_run<double, 3, 1>(c, ops, n, s) : enqueue_run<double, 3, 2>(c, ops, n, s);
}

static int enqueue_cycle_ops(evo_cycle *c, cudaStream_t s)
{
    // maximal runs of >= 2 consecutive statements on the small levels go out as ONE launch (evo_kernels_run.cuh)
    const size_t n = c->ops.size();
    for (size_t t = 0; t < n;) {
        size_t e = t;
        while (e < n && op_run_eligible(c, c->ops[e])) ++e;
        if (e - t >= 2) {
            EV(enqueue_fused_run(c, &c->ops[t], (int)(e - t), s));
            t = e;
        } else {
            const evo_op &op = c->ops[t];
            const evo_problem_desc &d = c->p->desc;
            // `RHS@(l-1) = R (f - A u)` followed by `SOL@(l-1) = 0`: one kernel writes both coarse fields
            if (op.code == EVO_OP_RESIDUAL_RESTRICT && t + 1 < n && c->ops[t + 1].code == EVO_OP_ZERO && c->ops[t + 1].level == op.level - 1 &&
                c->ops[t + 1].dst == EVO_BUF_SOL && d.kind == EVO_PROBLEM_LINEAR && d.scalar_words == 1 && c->zc_lo < 0 &&
                !c->coarse_sol_written && option(OPT_NO_ZERO_FUSE) == 0) {
                c->fuse_zero = true;
                const int rc = dispatch_op(c, op, s);
                const bool folded = !c->fuse_zero;
                c->fuse_zero = false;
                EV(rc);
                t += folded ? 2 : 1;
                continue;
            }
            EV(dispatch_op(c, op, s));
            ++t;
        }
    }
    return EVO_OK;
}

static int enqueue_cycle(evo_cycle *c, cudaStream_t s)
{
    evo_problem *p = c->p;
    EV(enqueue_cycle_ops(c, s));
    const size_t esz = sizeof(double) * p->words;
    for (int l = p->desc.min_level; l <= p->desc.max_level; ++l)
        for (int i = 0; i < p->desc.n_fields; ++i)
            if (c->lv[l].swapped[i]) {
                // current content lives in the buffer that was the slot at cycle start: copy back
                CU(cudaMemcpyAsync(c->lv[l].slot[i], c->lv[l].buf[EVO_BUF_SOL][i], (size_t)p->geom[l].total * esz,
                                   cudaMemcpyDeviceToDevice, s));
                bool cor_alias = c->lv[l].buf[EVO_BUF_COR][i] == c->lv[l].buf[EVO_BUF_SOL][i];
                std::swap(c->lv[l].buf[EVO_BUF_SOL][i], c->lv[l].slot[i]);
                if (cor_alias) c->lv[l].buf[EVO_BUF_COR][i] = c->lv[l].buf[EVO_BUF_SOL][i];
                c->lv[l].swapped[i] = false;
            }
    return EVO_OK;
}

// ------------------------------------------------------------------------------------------------
// cycle: validation + build
static int validate_ops(const evo_cycle *c)
{
    const evo_problem_desc &d = c->p->desc;
    for (size_t t = 0; t < c->ops.size(); ++t) {
        const evo_op &op = c->ops[t];
        int lo = d.min_level;
        if (op.code == EVO_OP_RESTRICT || op.code == EVO_OP_PROLONG_ADD || op.code == EVO_OP_PROLONG_SET ||
            op.code == EVO_OP_RESIDUAL_RESTRICT || op.code == EVO_OP_FAS_RESTRICT_SOL || op.code == EVO_OP_FAS_COARSE_RHS)
            lo = d.min_level + 1;
        if (op.level < lo || op.level > d.max_level) return fail(EVO_ERR_INVALID, "op %zu: level %d out of range", t, op.level);
        if (op.dst < 0 || op.dst >= EVO_BUF_COUNT || op.src < 0 || op.src >= EVO_BUF_COUNT)
            return fail(EVO_ERR_INVALID, "op %zu: invalid buffer", t);
        bool needs_op = op.code == EVO_OP_RESIDUAL || op.code == EVO_OP_RICHARDSON || op.code == EVO_OP_SMOOTH ||
                        op.code == EVO_OP_COARSE_SOLVE || op.code == EVO_OP_RESIDUAL_RESTRICT;
        if ((op.code == EVO_OP_FAS_RESTRICT_SOL || op.code == EVO_OP_FAS_COARSE_RHS || op.code == EVO_OP_FAS_SUB_APX) &&
            d.kind != EVO_PROBLEM_FAS)
            return fail(EVO_ERR_INVALID, "op %zu: FAS statement in a linear problem", t);
        if (needs_op && !c->has_sten[op.level]) return fail(EVO_ERR_INVALID, "op %zu: no operator for level %d", t, op.level);
        if (op.code == EVO_OP_SMOOTH && (op.n_unknowns < 1 || op.n_unknowns > EVO_MAX_UNKNOWNS))
            return fail(EVO_ERR_INVALID, "op %zu: local system size %d", t, op.n_unknowns);
    }
    if (!c->has_sten[d.max_level]) return fail(EVO_ERR_INVALID, "no operator for the finest level");
    return EVO_OK;
}

extern "C" int evo_cycle_destroy(evo_cycle *c);

// stencils, transfer weights and the descriptors of the smoothers as the fused-run kernels read them (device memory)
static int upload_run_tables(evo_cycle *c)
{
    evo_problem *p = c->p;
    if (p->desc.kind != EVO_PROBLEM_LINEAR || p->desc.scalar_words != 1) return EVO_OK;
    std::vector<SmoothParams> sp(std::max<size_t>(1, c->ops.size()));
    for (size_t t = 0; t < c->ops.size(); ++t) {
        const evo_op &op = c->ops[t];
        SmoothParams &q = sp[t];
        memset(&q, 0, sizeof(q));
        if (op.code != EVO_OP_SMOOTH) continue;
        q.nu = op.n_unknowns;
        q.omega = op.omega;
        q.color = -1;
        q.write_all = 0;
        for (int a = 0; a < q.nu && a < EVO_MAX_UNKNOWNS; ++a) {
            q.field[a] = op.unk_field[a];
            for (int d = 0; d < 3; ++d) q.off[a][d] = d < p->desc.dim ? op.unk_off[a][d] : 0;
        }
    }
    TransferW rp[2] = {p->R, p->P};
    cudaStream_t s = c->stream;
    CU(cudaMemcpyAsync(c->d_run_sten, c->sten, sizeof(OpSten) * EVO_MAX_LEVELS, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(c->d_run_sp, sp.data(), sizeof(SmoothParams) * sp.size(), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(c->d_run_rp, rp, sizeof(rp), cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));   // the sources are locals
    return EVO_OK;
}

extern "C" int evo_cycle_build(evo_problem *p, const evo_op *ops, int n_ops, const evo_level_operator *operators,
                               int n_operators, evo_cycle **out)
{
    if (!p || !out || (n_ops > 0 && !ops) || (n_operators > 0 && !operators)) return fail(EVO_ERR_INVALID, "null argument");
    CU(cudaSetDevice(p->desc.device));
    if (p->closed) return fail(EVO_ERR_INVALID, "problem already destroyed");
    evo_cycle *c = new evo_cycle();
    c->p = p;
    p->live_cycles++;
    // consecutive identical `solve locally` statements become one op with a repeat count (the kernels
    // fuse repeated sweeps; semantics are unchanged)
    for (int t = 0; t < n_ops; ++t) {
        evo_op op = ops[t];
        if (op.code == EVO_OP_SMOOTH && op.count < 1) op.count = 1;
        // (not for Helmholtz: the Robin boundary function is re-applied between two statements)
        if (!c->ops.empty() && op.code == EVO_OP_SMOOTH && op.kind == EVO_KIND_LINEAR && p->desc.kind != EVO_PROBLEM_HELMHOLTZ) {
            evo_op &prev = c->ops.back();
            evo_op a = prev, b = op;
            a.count = b.count = 0;
            if (prev.code == EVO_OP_SMOOTH && memcmp(&a, &b, sizeof(evo_op)) == 0) { prev.count += op.count; continue; }
        }
        c->ops.push_back(op);
    }
    c->slab = nullptr;
    c->graph = nullptr;
    c->exec = nullptr;
    c->stream = nullptr;
    c->h_state = nullptr;
    c->h_hist = nullptr;
    c->d_hist = nullptr;
    c->hist_cap = 0;
    c->use_while_graph = true;
    c->helm_graph = nullptr;
    c->helm_exec = nullptr;
    c->d_helm = nullptr;
    for (int v = 0; v < 9; ++v) c->helm[v] = nullptr;
    memset(c->has_sten, 0, sizeof(c->has_sten));
    memset(c->lv, 0, sizeof(c->lv));
    for (int t = 0; t < n_operators; ++t) {
        const int l = operators[t].level;
        if (l < p->desc.min_level || l > p->desc.max_level) { p->live_cycles--; delete c; return fail(EVO_ERR_INVALID, "operator level %d out of range", l); }
        OpSten &os = c->sten[l];
        memset(&os, 0, sizeof(os));
        for (int i = 0; i < p->desc.n_fields; ++i)
            for (int j = 0; j < p->desc.n_fields; ++j) {
                Sten &s = os.s[i][j];
                for (int q = 0; q < 27; ++q) {
                    double re = operators[t].coef[i][j][q][0], im = operators[t].coef[i][j][q][1];
                    if (re == 0.0 && im == 0.0) continue;
                    int oz = q / 9 - 1;
                    if (p->desc.dim == 2 && oz != 0) { p->live_cycles--; delete c; return fail(EVO_ERR_INVALID, "3-D stencil entry in a 2-D problem"); }
                    int k = s.nnz++;
                    s.ox[k] = (signed char)(q % 3 - 1); s.oy[k] = (signed char)((q / 3) % 3 - 1); s.oz[k] = (signed char)oz;
                    s.re[k] = re; s.im[k] = im;
                }
            }
        c->has_sten[l] = true;
    }
    // is RES@finest written before it is read by the cycle's statements?
    c->res_dead_on_entry = true;
    for (const evo_op &op : c->ops) {
        const int hi = p->desc.max_level;
        bool reads = ((op.code == EVO_OP_RESTRICT || op.code == EVO_OP_COPY) && op.level == hi && op.src == EVO_BUF_RES) ||
                     ((op.code == EVO_OP_PROLONG_ADD || op.code == EVO_OP_PROLONG_SET) && op.level - 1 == hi && op.src == EVO_BUF_RES) ||
                     (op.code == EVO_OP_FAS_COARSE_RHS && op.level == hi);
        bool writes = (op.code == EVO_OP_RESIDUAL && op.level == hi) ||
                      ((op.code == EVO_OP_COPY || op.code == EVO_OP_ZERO || op.code == EVO_OP_PROLONG_SET) && op.level == hi && op.dst == EVO_BUF_RES);
        if (reads) { c->res_dead_on_entry = false; break; }
        if (writes) break;
    }
    int rc = validate_ops(c);
    if (rc == EVO_OK) rc = allocate_cycle(c);
    if (rc != EVO_OK) { evo_cycle_destroy(c); return rc; }
    if (!p->res_pool.empty()) {
        evo_problem::CycleRes cr = p->res_pool.back();
        p->res_pool.pop_back();
        c->stream = cr.stream; c->ev0 = cr.ev0; c->ev1 = cr.ev1; c->h_state = cr.h_state;
        c->h_hist = cr.h_hist; c->d_hist = cr.d_hist; c->hist_cap = cr.hist_cap;
    } else {
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&c->ev0));
        CU(cudaEventCreate(&c->ev1));
        CU(cudaMallocHost(&c->h_state, sizeof(SolveState)));
    }
    rc = reset_cycle(c, c->stream);
    if (rc == EVO_OK) rc = upload_run_tables(c);
    if (rc != EVO_OK) { evo_cycle_destroy(c); return rc; }
    CU(cudaStreamSynchronize(c->stream));
    *out = c;
    return EVO_OK;
}

extern "C" int evo_cycle_destroy(evo_cycle *c)
{
    if (!c) return EVO_OK;
    cudaSetDevice(c->p->desc.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->exec) cudaGraphExecDestroy(c->exec);
    if (c->graph) cudaGraphDestroy(c->graph);
    if (c->helm_exec) cudaGraphExecDestroy(c->helm_exec);
    if (c->helm_graph) cudaGraphDestroy(c->helm_graph);
    if (c->slab) c->p->pool.push_back({c->slab, c->slab_cap});
    // keep the pool bounded: at most 512 slabs and a quarter of the device memory stay parked here
    size_t pooled = 0;
    for (auto &s : c->p->pool) pooled += s.second;
    while (!c->p->pool.empty() && (c->p->pool.size() > 512 || pooled > c->p->device_bytes / 4)) {
        pooled -= c->p->pool.front().second;
        cudaFree(c->p->pool.front().first);
        c->p->pool.erase(c->p->pool.begin());
    }
    if (c->stream && c->own_stream && c->h_state && !c->p->closed && c->p->res_pool.size() < 1024) {
        c->p->res_pool.push_back({c->stream, c->ev0, c->ev1, c->h_state, c->h_hist, c->d_hist, c->hist_cap});
    } else {
        if (c->d_hist) cudaFree(c->d_hist);
        if (c->h_state) cudaFreeHost(c->h_state);
        if (c->h_hist) cudaFreeHost(c->h_hist);
        if (c->stream) { cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); if (c->own_stream) cudaStreamDestroy(c->stream); }
    }
    evo_problem *p = c->p;
    delete c;
    if (--p->live_cycles == 0 && p->closed) free_problem(p);
    return EVO_OK;
}

extern "C" int evo_cycle_reset(evo_cycle *c)
{
    if (!c) return fail(EVO_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->p->desc.device));
    EV(reset_cycle(c, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return EVO_OK;
}

extern "C" int evo_cycle_apply(evo_cycle *c, int repeat)
{
    if (!c) return fail(EVO_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->p->desc.device));
    for (int r = 0; r < repeat; ++r) EV(enqueue_cycle(c, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return EVO_OK;
}

static int field_ptr(evo_cycle *c, int level, int buf, int field, void **out)
{
    const evo_problem_desc &d = c->p->desc;
    if (level < d.min_level || level > d.max_level || field < 0 || field >= d.n_fields || buf < 0 || buf >= EVO_BUF_COUNT)
        return fail(EVO_ERR_INVALID, "invalid level/buffer/field");
    *out = c->lv[level].buf[buf][field];
    if (!*out) return fail(EVO_ERR_INVALID, "buffer not allocated for this problem kind");
    return EVO_OK;
}

extern "C" int evo_cycle_get_field(evo_cycle *c, int level, int buf, int field, double *host, size_t n_doubles)
{
    if (!c || !host) return fail(EVO_ERR_INVALID, "null argument");
    void *dev;
    EV(field_ptr(c, level, buf, field, &dev));
    const Geom &g = c->p->geom[level];
    if (n_doubles != (size_t)g.n * g.n * (g.dim == 3 ? g.n : 1) * c->p->words) return fail(EVO_ERR_INVALID, "field size mismatch");
    CU(cudaSetDevice(c->p->desc.device));
    return copy_field(g, c->p->words, dev, nullptr, host, c->stream);
}

extern "C" int evo_cycle_set_field(evo_cycle *c, int level, int buf, int field, const double *host, size_t n_doubles)
{
    if (!c || !host) return fail(EVO_ERR_INVALID, "null argument");
    c->pristine = false;
    void *dev;
    EV(field_ptr(c, level, buf, field, &dev));
    if (level == c->p->desc.max_level && buf == EVO_BUF_RHS && c->p->desc.kind != EVO_PROBLEM_HELMHOLTZ)
        if (c->p->desc.kind != EVO_PROBLEM_HELMHOLTZ)
            return fail(EVO_ERR_UNSUPPORTED, "the finest right-hand side is shared by all cycles: use evo_problem_set_field");
    const Geom &g = c->p->geom[level];
    if (n_doubles != (size_t)g.n * g.n * (g.dim == 3 ? g.n : 1) * c->p->words) return fail(EVO_ERR_INVALID, "field size mismatch");
    CU(cudaSetDevice(c->p->desc.device));
    EV(copy_field(g, c->p->words, dev, host, nullptr, c->stream));
    if (buf == EVO_BUF_SOL && level < c->p->desc.max_level && !c->coarse_sol_written) {
        // the caller may have put values on the boundary layer of a correction level: from now on `SOL = 0` clears the
        // whole array again (solver graphs captured before are rebuilt)
        c->coarse_sol_written = true;
        if (c->exec) { cudaGraphExecDestroy(c->exec); c->exec = nullptr; }
        if (c->graph) { cudaGraphDestroy(c->graph); c->graph = nullptr; }
    }
    if (buf == EVO_BUF_SOL && c->lv[level].slot[field])  // keep the boundary invariant of the jacobi slot
        EV(copy_field(g, c->p->words, c->lv[level].slot[field], host, nullptr, c->stream));
    return EVO_OK;
}

extern "C" int evo_cycle_residual_norm(evo_cycle *c, double *norm)
{
    if (!c || !norm) return fail(EVO_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->p->desc.device));
    EV(dispatch_residual_norm(c, c->stream, true));
    CU(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(SolveState), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *norm = sqrt(c->h_state->sum);
    return EVO_OK;
}

extern "C" int evo_cycle_set_stream(evo_cycle *c, void *cuda_stream)
{
    if (!c) return fail(EVO_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->p->desc.device));
    if (c->stream && c->own_stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return EVO_OK;
}

extern "C" int evo_cycle_exec_ops(evo_cycle *c, const evo_op *ops, int n_ops, int zc_lo, int zc_hi)
{
    if (!c || (n_ops > 0 && !ops)) return fail(EVO_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->p->desc.device));
    std::vector<evo_op> saved = c->ops;
    c->ops.assign(ops, ops + n_ops);
    int rc = validate_ops(c);
    c->ops = saved;
    EV(rc);
    c->zc_lo = zc_lo; c->zc_hi = zc_hi;
    for (int t = 0; t < n_ops && rc == EVO_OK; ++t) rc = dispatch_op(c, ops[t], c->stream);
    c->zc_lo = c->zc_hi = -1;
    return rc;
}

extern "C" int evo_cycle_buffer(evo_cycle *c, int level, int buf, int field, void **device_ptr)
{
    if (!c || !device_ptr) return fail(EVO_ERR_INVALID, "null argument");
    if (buf == EVO_BUF_NEXT) {   // the [next] slot of SOL (target of out-of-place statements)
        const evo_problem_desc &d = c->p->desc;
        if (level < d.min_level || level > d.max_level || field < 0 || field >= d.n_fields || !c->lv[level].slot[field])
            return fail(EVO_ERR_INVALID, "no [next] slot on level %d", level);
        *device_ptr = c->lv[level].slot[field];
        return EVO_OK;
    }
    return field_ptr(c, level, buf, field, device_ptr);
}

// one statement on the local planes [z_lo, z_hi] only; with EVO_PART_NO_SWAP an out-of-place smoother leaves SOL and
// its [next] slot unexchanged, so that further parts of the same statement (and the halo exchange of the planes
// already written) can follow; the last part is executed without the flag
extern "C" int evo_cycle_exec_part(evo_cycle *c, const evo_op *op, int z_lo, int z_hi, int flags)
{
    if (!c || !op) return fail(EVO_ERR_INVALID, "null argument");
    if (op->level < c->p->desc.min_level || op->level > c->p->desc.max_level) return fail(EVO_ERR_INVALID, "invalid level");
    CU(cudaSetDevice(c->p->desc.device));
    c->zc_lo = z_lo; c->zc_hi = z_hi;
    c->part_no_swap = (flags & EVO_PART_NO_SWAP) != 0;
    int rc = z_hi >= z_lo ? dispatch_op(c, *op, c->stream) : EVO_OK;
    if (z_hi < z_lo && !c->part_no_swap) {   // empty last part: only the exchange of the slots
        c->zc_lo = 1; c->zc_hi = 0;
        rc = dispatch_op(c, *op, c->stream);
    }
    c->zc_lo = c->zc_hi = -1;
    c->part_no_swap = false;
    return rc;
}

extern "C" int evo_cycle_residual_plane_sums(evo_cycle *c, double **device_sums, int *count)
{
    if (!c || !device_sums || !count) return fail(EVO_ERR_INVALID, "null argument");
    const evo_problem_desc &d = c->p->desc;
    if (d.dim != 3 || d.n_fields != 1 || d.scalar_words != 1) return fail(EVO_ERR_UNSUPPORTED, "3-D real scalar only");
    CU(cudaSetDevice(d.device));
    const int l = d.max_level;
    const Geom &g = c->p->geom[l];
    auto u = fields_of<double>(c->lv[l].buf[EVO_BUF_SOL], 1), f = fields_of<double>(c->lv[l].buf[EVO_BUF_RHS], 1),
         r = fields_of<double>(c->lv[l].buf[EVO_BUF_RES], 1);
    if (!star::try_residual_norm<double, 3, 1>(g, c->sten[l], u, f, r, c->d_partials, true, c->stream))
        return fail(EVO_ERR_UNSUPPORTED, "residual fast path not applicable");
    const int ni = g.n - 2, planes = g.zhi - g.zlo + 1;
    double *psums = c->d_partials + (size_t)ni * planes;
    // one warp per owned plane: canonical vecsum over its rows
    k_reduce_planes<<<(unsigned)((planes + 7) / 8), 256, 0, c->stream>>>(c->d_partials, planes, ni, psums);
    CU(cudaGetLastError());
    *device_sums = psums;
    *count = planes;
    return EVO_OK;
}

__global__ void k_vecsum_out(const double *vals, int m, double *out)
{
    double s = warp_vecsum(vals, m);
    if (threadIdx.x == 0) *out = s;
}

extern "C" int evo_cycle_vecsum(evo_cycle *c, const double *device_vals, int m, double *out)
{
    if (!c || !device_vals || !out || m < 0) return fail(EVO_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->p->desc.device));
    k_vecsum_out<<<1, 32, 0, c->stream>>>(device_vals, m, &c->d_state->sum);
    CU(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(SolveState), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *out = c->h_state->sum;
    return EVO_OK;
}

// stream-ordered variant for captured execution: the sum stays on the device (SolveState::sum) until
// evo_cycle_read_sum copies it out
extern "C" int evo_cycle_vecsum_async(evo_cycle *c, const double *device_vals, int m)
{
    if (!c || !device_vals || m < 0) return fail(EVO_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->p->desc.device));
    k_vecsum_out<<<1, 32, 0, c->stream>>>(device_vals, m, &c->d_state->sum);
    CU(cudaGetLastError());
    return EVO_OK;
}

extern "C" int evo_cycle_read_sum(evo_cycle *c, double *out)
{
    if (!c || !out) return fail(EVO_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->p->desc.device));
    CU(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(SolveState), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *out = c->h_state->sum;
    return EVO_OK;
}

// exchange SOL and its [next] slot on the host side (after replaying a captured cycle an odd number of times the
// data lives in the other buffer)
extern "C" int evo_cycle_swap_slots(evo_cycle *c, int level)
{
    if (!c || level < c->p->desc.min_level || level > c->p->desc.max_level) return fail(EVO_ERR_INVALID, "invalid argument");
    for (int i = 0; i < c->p->desc.n_fields; ++i) {
        if (!c->lv[level].slot[i]) return fail(EVO_ERR_INVALID, "level %d has no [next] slot", level);
        bool cor_alias = c->lv[level].buf[EVO_BUF_COR][i] == c->lv[level].buf[EVO_BUF_SOL][i];
        std::swap(c->lv[level].buf[EVO_BUF_SOL][i], c->lv[level].slot[i]);
        if (cor_alias) c->lv[level].buf[EVO_BUF_COR][i] = c->lv[level].buf[EVO_BUF_SOL][i];
        c->lv[level].swapped[i] = !c->lv[level].swapped[i];
    }
    return EVO_OK;
}

extern "C" int evo_cycle_profile_op(evo_cycle *c, const evo_op *op, int repeat, double *ms_per_exec, int64_t *launches_per_exec)
{
    if (!c || !op || repeat < 1) return fail(EVO_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->p->desc.device));
    // validate like evo_cycle_build does
    std::vector<evo_op> saved = c->ops;
    c->ops.assign(1, *op);
    int rc = validate_ops(c);
    c->ops = saved;
    EV(rc);
    if (op->code == EVO_OP_SMOOTH && op->mode == EVO_SMOOTH_JACOBI && !c->lv[op->level].slot[0])
        return fail(EVO_ERR_INVALID, "the cycle was built without a jacobi slot on level %d", op->level);
    cudaStream_t s = c->stream;
    EV(dispatch_op(c, *op, s));  // warm-up (lazy attribute settings, first-touch of the kernels)
    CU(cudaStreamSynchronize(s));
    // the statement `repeat` times as kernel nodes of ONE graph: what it costs inside a solver graph (an eager loop of
    // tiny launches would measure the host's launch rate instead)
    c->launch_counter = 0;
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int rc2 = EVO_OK;
    for (int r = 0; r < repeat && rc2 == EVO_OK; ++r) rc2 = dispatch_op(c, *op, s);
    cudaError_t ce = cudaStreamEndCapture(s, &g);
    if (rc2 != EVO_OK || ce != cudaSuccess) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        if (rc2 != EVO_OK) return rc2;
        return fail(EVO_ERR_CUDA, "profile capture: %s", cudaGetErrorString(ce));
    }
    ce = cudaGraphInstantiate(&ge, g, 0);
    float ms = 0.f;
    if (ce == cudaSuccess) ce = cudaGraphLaunch(ge, s);          // warm-up of the graph itself
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    if (ce == cudaSuccess) ce = cudaEventRecord(c->ev0, s);
    if (ce == cudaSuccess) ce = cudaGraphLaunch(ge, s);
    if (ce == cudaSuccess) ce = cudaEventRecord(c->ev1, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    if (ge) cudaGraphExecDestroy(ge);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) return fail(EVO_ERR_CUDA, "profile graph: %s", cudaGetErrorString(ce));
    if (ms_per_exec) *ms_per_exec = (double)ms / repeat;
    if (launches_per_exec) *launches_per_exec = c->launch_counter / repeat;
    // restore the canonical jacobi slot assignment
    for (int l = c->p->desc.min_level; l <= c->p->desc.max_level; ++l)
        for (int i = 0; i < c->p->desc.n_fields; ++i)
            if (c->lv[l].swapped[i]) {
                bool cor_alias = c->lv[l].buf[EVO_BUF_COR][i] == c->lv[l].buf[EVO_BUF_SOL][i];
                std::swap(c->lv[l].buf[EVO_BUF_SOL][i], c->lv[l].slot[i]);
                if (cor_alias) c->lv[l].buf[EVO_BUF_COR][i] = c->lv[l].buf[EVO_BUF_SOL][i];
                c->lv[l].swapped[i] = false;
            }
    return EVO_OK;
}

// ------------------------------------------------------------------------------------------------
// solve: the generated solver's outer loop
__global__ void k_set_while_condition(cudaGraphConditionalHandle handle, const SolveState *st)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) cudaGraphSetConditional(handle, st->done ? 0u : 1u);
}

// after the loop: an odd number of cycles leaves the solution in the [next] slots
__global__ void k_set_odd_condition(cudaGraphConditionalHandle h, const SolveState *st)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) cudaGraphSetConditional(h, (unsigned)(st->it & 1));
}

static int graph_leaves(cudaGraph_t g, std::vector<cudaGraphNode_t> &leaves)
{
    size_t n_nodes = 0;
    CU(cudaGraphGetNodes(g, nullptr, &n_nodes));
    std::vector<cudaGraphNode_t> nodes(n_nodes);
    CU(cudaGraphGetNodes(g, nodes.data(), &n_nodes));
    leaves.clear();
    for (cudaGraphNode_t nd : nodes) {
        size_t nd_dep = 0;
        CU(cudaGraphNodeGetDependentNodes(nd, nullptr, &nd_dep));
        if (nd_dep == 0) leaves.push_back(nd);
    }
    return EVO_OK;
}

static int ensure_hist(evo_cycle *c, int max_iters)
{
    if (c->hist_cap >= max_iters + 1) return EVO_OK;
    if (c->d_hist) cudaFree(c->d_hist);
    if (c->h_hist) cudaFreeHost(c->h_hist);
    c->hist_cap = max_iters + 1;
    CU(cudaMalloc(&c->d_hist, sizeof(double) * c->hist_cap));
    CU(cudaMallocHost(&c->h_hist, sizeof(double) * c->hist_cap));
    return EVO_OK;
}

// Error paths of the graph builders: a failing call must not leave the stream in capture mode or leak the graph.
struct CaptureGuard {       // ends an open capture (the graph of a capture-to-graph belongs to its parent)
    cudaStream_t s;
    bool open = false, owns_graph = false;
    explicit CaptureGuard(cudaStream_t st) : s(st) {}
    ~CaptureGuard()
    {
        if (!open) return;
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(s, &g);
        if (g && owns_graph) cudaGraphDestroy(g);
        cudaGetLastError();
    }
};
struct GraphGuard {         // destroys a graph under construction unless release()d
    cudaGraph_t g = nullptr;
    ~GraphGuard() { if (g) cudaGraphDestroy(g); }
    cudaGraph_t release() { cudaGraph_t r = g; g = nullptr; return r; }
};

// graph = [res0 + norm + update(0)] -> WHILE(!done) { cycle ops; residual + norm; update(1); set condition }
static int build_solver_graph(evo_cycle *c, double tol, int max_iters)
{
    if (c->exec && c->graph_tol == tol && c->graph_max_iters == max_iters) return EVO_OK;
    if (c->exec) { cudaGraphExecDestroy(c->exec); c->exec = nullptr; }
    if (c->graph) { cudaGraphDestroy(c->graph); c->graph = nullptr; }
    cudaStream_t s = c->stream;
    // prologue captured into the main graph
    c->launch_counter = 0;
    CaptureGuard cap(s);
    GraphGuard gg;
    CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    cap.open = true; cap.owns_graph = true;
    int rc = dispatch_residual_norm(c, s);
    if (rc == EVO_OK) {
        k_outer_update<<<1, 32, 0, s>>>(c->d_state, c->d_hist, tol, max_iters, 0);
        c->launch_counter++;
    }
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(s, &g);
    cap.open = false;
    gg.g = g;
    if (rc != EVO_OK) return rc;
    if (e != cudaSuccess) return fail(EVO_ERR_CUDA, "prologue capture: %s", cudaGetErrorString(e));
    c->kernels_prologue = c->launch_counter;
    // find the leaf of the prologue to hang the while node on
    size_t n_nodes = 0;
    CU(cudaGraphGetNodes(g, nullptr, &n_nodes));
    std::vector<cudaGraphNode_t> nodes(n_nodes);
    CU(cudaGraphGetNodes(g, nodes.data(), &n_nodes));
    std::vector<cudaGraphNode_t> leaves;
    for (cudaGraphNode_t nd : nodes) {
        size_t nd_dep = 0;
        CU(cudaGraphNodeGetDependentNodes(nd, nullptr, &nd_dep));
        if (nd_dep == 0) leaves.push_back(nd);
    }
    cudaGraphConditionalHandle handle;
    CU(cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    cudaGraphNode_t wnode;
    // guard: the while body must not run when the prologue already set done (max_iters == 0 / bad)
    // -> a tiny kernel node after the prologue sets the condition from the state
    cudaGraphNode_t cond_init;
    {
        cudaKernelNodeParams kp;
        memset(&kp, 0, sizeof(kp));
        void *args[2] = {(void *)&handle, (void *)&c->d_state};
        kp.func = (void *)k_set_while_condition;
        kp.gridDim = dim3(1);
        kp.blockDim = dim3(32);
        kp.kernelParams = args;
        CU(cudaGraphAddKernelNode(&cond_init, g, leaves.data(), leaves.size(), &kp));
        c->kernels_prologue++;
    }
    CU(cudaGraphAddNode(&wnode, g, &cond_init, 1, &cp));
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    c->launch_counter = 0;
    CU(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    cap.open = true; cap.owns_graph = false;
    // Out-of-place statements (jacobi slots, the streaming RB-GS kernel) swap SOL and its [next] slot.  A cycle
    // with an odd number of swaps on a level ends in the other buffer: instead of copying it back every cycle
    // (2 x 8 B/DOF of pure overhead) the loop body holds TWO cycles -- the second one, inside an IF node, runs
    // with the roles of the buffers exchanged and ends in the canonical assignment again.
    rc = enqueue_cycle_ops(c, s);
    bool odd = false;
    for (int l = c->p->desc.min_level; l <= c->p->desc.max_level; ++l)
        for (int i = 0; i < c->p->desc.n_fields; ++i) {
            c->odd_swap[l][i] = c->lv[l].swapped[i];
            odd = odd || c->lv[l].swapped[i];
        }
    const bool no_pingpong = option(OPT_NO_PINGPONG) != 0;
    c->pingpong = odd && !no_pingpong && rc == EVO_OK;
    if (rc == EVO_OK && !c->pingpong) {
        // restore the canonical assignment by copying (no-op when nothing is swapped)
        std::vector<evo_op> none;
        std::swap(none, c->ops);
        rc = enqueue_cycle(c, s);
        std::swap(none, c->ops);
    }
    cudaGraphConditionalHandle h_if = 0;
    if (rc == EVO_OK && c->pingpong) CU(cudaGraphConditionalHandleCreate(&h_if, g, 0, cudaGraphCondAssignDefault));
    // the last reduction of the norm also updates the loop state and sets the loop's condition(s)
    struct FinishOff { evo_cycle *c; ~FinishOff() { c->fin.on = false; } } finish_off{c};
    c->fin.on = true; c->fin.tol = tol; c->fin.max_iters = max_iters; c->fin.mode = 1;
    c->fin.n_handles = c->pingpong ? 2 : 1; c->fin.h[0] = handle; c->fin.h[1] = h_if;
    if (rc == EVO_OK) rc = dispatch_residual_norm(c, s);
    c->fin.on = false;
    c->kernels_per_cycle = c->launch_counter;
    e = cudaStreamEndCapture(s, nullptr);
    cap.open = false;
    if (rc != EVO_OK) return rc;
    if (e != cudaSuccess) return fail(EVO_ERR_CUDA, "body capture: %s", cudaGetErrorString(e));
    if (c->pingpong) {
        std::vector<cudaGraphNode_t> bl;
        EV(graph_leaves(body, bl));
        cudaGraphNodeParams ip = {cudaGraphNodeTypeConditional};
        ip.type = cudaGraphNodeTypeConditional;
        ip.conditional.handle = h_if;
        ip.conditional.type = cudaGraphCondTypeIf;
        ip.conditional.size = 1;
        cudaGraphNode_t inode;
        CU(cudaGraphAddNode(&inode, body, bl.data(), bl.size(), &ip));
        cudaGraph_t second = ip.conditional.phGraph_out[0];
        CU(cudaStreamBeginCaptureToGraph(s, second, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        cap.open = true; cap.owns_graph = false;
        rc = enqueue_cycle_ops(c, s);
        bool still = false;
        for (int l = c->p->desc.min_level; l <= c->p->desc.max_level; ++l)
            for (int i = 0; i < c->p->desc.n_fields; ++i) still = still || c->lv[l].swapped[i];
        if (rc == EVO_OK && still) rc = fail(EVO_ERR_INVALID, "ping-pong body did not return to the canonical slots");
        c->fin.on = true; c->fin.n_handles = 1; c->fin.h[0] = handle; c->fin.h[1] = 0;
        if (rc == EVO_OK) rc = dispatch_residual_norm(c, s);
        c->fin.on = false;
        e = cudaStreamEndCapture(s, nullptr);
        cap.open = false;
        if (rc != EVO_OK) return rc;
        if (e != cudaSuccess) return fail(EVO_ERR_CUDA, "second body capture: %s", cudaGetErrorString(e));
        // epilogue: after an odd number of cycles copy the solution back into the canonical buffers (once per solve)
        cudaGraphConditionalHandle h_odd;
        CU(cudaGraphConditionalHandleCreate(&h_odd, g, 0, cudaGraphCondAssignDefault));
        cudaGraphNode_t seto;
        {
            cudaKernelNodeParams kp;
            memset(&kp, 0, sizeof(kp));
            void *args[2] = {(void *)&h_odd, (void *)&c->d_state};
            kp.func = (void *)k_set_odd_condition;
            kp.gridDim = dim3(1);
            kp.blockDim = dim3(32);
            kp.kernelParams = args;
            CU(cudaGraphAddKernelNode(&seto, g, &wnode, 1, &kp));
        }
        cudaGraphNodeParams op = {cudaGraphNodeTypeConditional};
        op.type = cudaGraphNodeTypeConditional;
        op.conditional.handle = h_odd;
        op.conditional.type = cudaGraphCondTypeIf;
        op.conditional.size = 1;
        cudaGraphNode_t onode;
        CU(cudaGraphAddNode(&onode, g, &seto, 1, &op));
        cudaGraph_t fix = op.conditional.phGraph_out[0];
        CU(cudaStreamBeginCaptureToGraph(s, fix, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        cap.open = true; cap.owns_graph = false;
        const size_t esz = sizeof(double) * c->p->words;
        cudaError_t ce = cudaSuccess;
        {
            // the data of an odd solve lives where the pointers were after the FIRST cycle: slot <-> SOL exchanged
            for (int l = c->p->desc.min_level; l <= c->p->desc.max_level && ce == cudaSuccess; ++l)
                for (int i = 0; i < c->p->desc.n_fields && ce == cudaSuccess; ++i)
                    if (c->odd_swap[l][i])
                        ce = cudaMemcpyAsync(c->lv[l].buf[EVO_BUF_SOL][i], c->lv[l].slot[i], (size_t)c->p->geom[l].total * esz,
                                             cudaMemcpyDeviceToDevice, s);
        }
        e = cudaStreamEndCapture(s, nullptr);
        cap.open = false;
        if (ce != cudaSuccess || e != cudaSuccess) return fail(EVO_ERR_CUDA, "epilogue capture failed");
    }
    // (kernels_per_cycle was taken after the FIRST body: the second half of a ping-pong body is another cycle, not more
    // launches per cycle)
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, g, 0);
    if (e != cudaSuccess) return fail(EVO_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    c->graph = gg.release();
    c->exec = exec;
    c->graph_tol = tol;
    c->graph_max_iters = max_iters;
    return EVO_OK;
}

static __global__ void k_set_cap(SolveState *st, int cap) { st->cap = cap; }
static __global__ void k_set_timeout(SolveState *st, unsigned long long ns) { st->timeout_ns = ns; }
static __global__ void k_set_timeout_helm(helm::HelmState *st, unsigned long long ns) { st->timeout_ns = ns; }

// enqueue one complete solve on the cycle's stream (no host synchronisation)
static int enqueue_solve(evo_cycle *c, const evo_solve_params *prm)
{
    cudaStream_t s = c->stream;
    if (!(prm->flags & EVO_SOLVE_KEEP_STATE) && !c->pristine) EV(reset_cycle(c, s));   // a fresh cycle is already reset
    c->pristine = false;
    if (prm->timeout_ms > 0) k_set_timeout<<<1, 1, 0, s>>>(c->d_state, (unsigned long long)prm->timeout_ms * 1000000ull);
    CU(cudaEventRecord(c->ev0, s));
    if (prm->flags & EVO_SOLVE_NO_GRAPH) {
        // debugging path: direct launches, host-side loop with one synchronisation per iteration
        c->launch_counter = 0;
        EV(dispatch_residual_norm(c, s));
        k_outer_update<<<1, 32, 0, s>>>(c->d_state, c->d_hist, prm->tol, prm->max_iters, 0);
        c->kernels_prologue = c->launch_counter + 1;
        for (int it = 0; it < prm->max_iters; ++it) {
            CU(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(SolveState), cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
            if (c->h_state->done) break;
            c->launch_counter = 0;
            EV(enqueue_cycle(c, s));
            EV(dispatch_residual_norm(c, s));
            k_outer_update<<<1, 32, 0, s>>>(c->d_state, c->d_hist, prm->tol, prm->max_iters, 1);
            c->kernels_per_cycle = c->launch_counter + 1;
        }
    } else {
        CU(cudaGraphLaunch(c->exec, s));
    }
    CU(cudaEventRecord(c->ev1, s));
    CU(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(SolveState), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(c->h_hist, c->d_hist, sizeof(double) * (prm->max_iters + 1), cudaMemcpyDeviceToHost, s));
    return EVO_OK;
}

static int collect_solve(evo_cycle *c, const evo_solve_params *prm, evo_solve_result *res, double *res_hist, float *ms)
{
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    const SolveState &st = *c->h_state;
    res->status = st.bad ? 1 : (st.timed_out ? 2 : 0);
    res->iterations = st.it;
    res->initial_residual = st.res0;
    res->final_residual = st.res;
    res->kernel_launches = c->kernels_prologue + c->kernels_per_cycle * (int64_t)st.it;
    if (res_hist) memcpy(res_hist, c->h_hist, sizeof(double) * (st.it + 1));
    return EVO_OK;
}

static int prepare_solve(evo_cycle *c, const evo_solve_params *prm)
{
    if (prm->max_iters < 0 || prm->max_iters > 1000000) return fail(EVO_ERR_INVALID, "max_iters out of range");
    CU(cudaSetDevice(c->p->desc.device));
    EV(ensure_hist(c, prm->max_iters));
    if (!(prm->flags & EVO_SOLVE_NO_GRAPH)) EV(build_solver_graph(c, prm->tol, prm->max_iters));
    return EVO_OK;
}

extern "C" int evo_cycle_solve(evo_cycle *c, const evo_solve_params *prm, evo_solve_result *res, double *res_hist)
{
    if (!c || !prm || !res) return fail(EVO_ERR_INVALID, "null argument");
    EV(prepare_solve(c, prm));
    const int samples = prm->samples > 0 ? prm->samples : 1;
    std::vector<float> times;
    memset(res, 0, sizeof(*res));
    for (int sidx = 0; sidx < samples; ++sidx) {
        float ms = 0.f;
        EV(enqueue_solve(c, prm));
        EV(collect_solve(c, prm, res, res_hist, &ms));
        times.push_back(ms);
    }
    std::sort(times.begin(), times.end());
    res->time_ms = times[times.size() / 2];
    res->time_ms_min = times[0];
    return EVO_OK;
}

// ------------------------------------------------------------------------------------------------
// Helmholtz: outer PreconditionedBiCGStab@finest (exa3:144-200) with the cycle as preconditioner, one CUDA
// graph: prologue + WHILE(!done) { one BiCGStab iteration = 2 cycle applications }
static int helm_dot(evo_cycle *c, const cplx *a, const cplx *b, int slot, cudaStream_t s)
{
    const Geom &g = c->p->geom[c->p->desc.max_level];
    const int ni = g.n - 2;
    helm::k2_cdot_rows<<<(ni + 3) / 4, 128, 0, s>>>(g, a, b, (cplx *)c->helm[8]);
    helm::k2_cdot_final<<<1, 32, 0, s>>>((const cplx *)c->helm[8], ni, c->d_helm, slot);
    c->launch_counter += 2;
    CU(cudaGetLastError());
    return EVO_OK;
}

static int helm_precondition(evo_cycle *c, const cplx *src, cudaStream_t s)
{
    // u = 0; f = src; gen_mgCycle()
    const int hi = c->p->desc.max_level;
    const Geom &g = c->p->geom[hi];
    CU(cudaMemsetAsync(c->lv[hi].buf[EVO_BUF_SOL][0], 0, (size_t)g.total * sizeof(cplx), s));
    helm::k2_helm_vec<<<row_grid(g), BX, 0, s>>>(g, 4, c->d_helm, (cplx *)c->lv[hi].buf[EVO_BUF_RHS][0], nullptr, src, nullptr, nullptr, nullptr);
    c->launch_counter++;
    CU(cudaGetLastError());
    return enqueue_cycle(c, s);
}

static int helm_iteration(evo_cycle *c, double tol, int max_iters, cudaStream_t s)
{
    const int hi = c->p->desc.max_level;
    const Geom &g = c->p->geom[hi];
    cplx *x = (cplx *)c->helm[0], *r = (cplx *)c->helm[1], *p = (cplx *)c->helm[2], *ap = (cplx *)c->helm[3],
         *sv = (cplx *)c->helm[4], *t = (cplx *)c->helm[5], *h = (cplx *)c->helm[6], *rh = (cplx *)c->helm[7];
    const dim3 grid = row_grid(g);
    EV(helm_dot(c, rh, r, 0, s));
    helm::k_helm_scalar<<<1, 32, 0, s>>>(c->d_helm, c->d_hist, tol, max_iters, 1);
    helm::k2_helm_vec<<<grid, BX, 0, s>>>(g, 0, c->d_helm, p, nullptr, r, ap, nullptr, nullptr);
    c->launch_counter += 2;
    EV(helm_precondition(c, p, s));
    cplx *u = (cplx *)c->lv[hi].buf[EVO_BUF_SOL][0];
    helm::k2_apply_op<<<grid, BX, 0, s>>>(g, c->helm_A, u, ap, nullptr);
    c->launch_counter++;
    EV(helm_dot(c, rh, ap, 1, s));
    helm::k_helm_scalar<<<1, 32, 0, s>>>(c->d_helm, c->d_hist, tol, max_iters, 2);
    helm::k2_helm_vec<<<grid, BX, 0, s>>>(g, 1, c->d_helm, h, sv, x, u, r, ap);
    c->launch_counter += 2;
    EV(helm_precondition(c, sv, s));
    u = (cplx *)c->lv[hi].buf[EVO_BUF_SOL][0];
    helm::k2_apply_op<<<grid, BX, 0, s>>>(g, c->helm_A, u, t, nullptr);
    c->launch_counter++;
    EV(helm_dot(c, t, sv, 2, s));
    EV(helm_dot(c, t, t, 3, s));
    helm::k_helm_scalar<<<1, 32, 0, s>>>(c->d_helm, c->d_hist, tol, max_iters, 3);
    helm::k2_helm_vec<<<grid, BX, 0, s>>>(g, 2, c->d_helm, x, nullptr, h, u, nullptr, nullptr);
    c->launch_counter += 2;
    EV(helm_bc(c, hi, x, s));
    helm::k2_helm_vec<<<grid, BX, 0, s>>>(g, 3, c->d_helm, r, nullptr, sv, t, nullptr, nullptr);
    c->launch_counter++;
    EV(helm_dot(c, r, r, 0, s));
    helm::k_helm_scalar<<<1, 32, 0, s>>>(c->d_helm, c->d_hist, tol, max_iters, 4);
    c->launch_counter++;
    CU(cudaGetLastError());
    return EVO_OK;
}

static int helm_prologue(evo_cycle *c, double tol, int max_iters, cudaStream_t s)
{
    const int hi = c->p->desc.max_level;
    const Geom &g = c->p->geom[hi];
    const size_t bytes = (size_t)g.total * sizeof(cplx);
    cplx *x = (cplx *)c->helm[0], *r = (cplx *)c->helm[1], *rh = (cplx *)c->helm[7];
    CU(cudaMemcpyAsync(x, c->p->init_sol[0], bytes, cudaMemcpyDeviceToDevice, s));   // Solution = 0 (+ boundary function)
    for (int v = 2; v < 7; ++v) CU(cudaMemsetAsync(c->helm[v], 0, bytes, s));
    CU(cudaMemsetAsync(r, 0, bytes, s));
    EV(helm_bc(c, hi, x, s));
    helm::k2_apply_op<<<row_grid(g), BX, 0, s>>>(g, c->helm_A, x, r, (const cplx *)c->p->rhs0[0]);   // Residual = RHS - A Solution
    c->launch_counter++;
    EV(helm_dot(c, r, r, 0, s));
    helm::k_helm_scalar<<<1, 32, 0, s>>>(c->d_helm, c->d_hist, tol, max_iters, 0);
    c->launch_counter++;
    CU(cudaMemcpyAsync(rh, r, bytes, cudaMemcpyDeviceToDevice, s));
    CU(cudaGetLastError());
    return EVO_OK;
}

static int build_helm_graph(evo_cycle *c, double tol, int max_iters)
{
    // the un-shifted operator is a by-value kernel argument of the captured graph: a different A needs a new graph
    if (c->helm_exec && c->helm_tol == tol && c->helm_max_iters == max_iters &&
        memcmp(&c->helm_A_graph, &c->helm_A, sizeof(OpSten)) == 0)
        return EVO_OK;
    if (c->helm_exec) { cudaGraphExecDestroy(c->helm_exec); c->helm_exec = nullptr; }
    if (c->helm_graph) { cudaGraphDestroy(c->helm_graph); c->helm_graph = nullptr; }
    cudaStream_t s = c->stream;
    c->launch_counter = 0;
    CaptureGuard cap(s);
    GraphGuard gg;
    CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    cap.open = true; cap.owns_graph = true;
    int rc = helm_prologue(c, tol, max_iters, s);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(s, &g);
    cap.open = false;
    gg.g = g;
    if (rc != EVO_OK) return rc;
    if (e != cudaSuccess) return fail(EVO_ERR_CUDA, "helmholtz prologue capture: %s", cudaGetErrorString(e));
    c->kernels_prologue = c->launch_counter;
    size_t n_nodes = 0;
    CU(cudaGraphGetNodes(g, nullptr, &n_nodes));
    std::vector<cudaGraphNode_t> nodes(n_nodes), leaves;
    CU(cudaGraphGetNodes(g, nodes.data(), &n_nodes));
    for (cudaGraphNode_t nd : nodes) {
        size_t nd_dep = 0;
        CU(cudaGraphNodeGetDependentNodes(nd, nullptr, &nd_dep));
        if (nd_dep == 0) leaves.push_back(nd);
    }
    cudaGraphConditionalHandle handle;
    CU(cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNode_t cond_init, wnode;
    {
        cudaKernelNodeParams kp;
        memset(&kp, 0, sizeof(kp));
        void *args[2] = {(void *)&handle, (void *)&c->d_helm};
        kp.func = (void *)helm::k_set_while_condition_helm;
        kp.gridDim = dim3(1);
        kp.blockDim = dim3(32);
        kp.kernelParams = args;
        CU(cudaGraphAddKernelNode(&cond_init, g, leaves.data(), leaves.size(), &kp));
        c->kernels_prologue++;
    }
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    CU(cudaGraphAddNode(&wnode, g, &cond_init, 1, &cp));
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    c->launch_counter = 0;
    CU(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    cap.open = true; cap.owns_graph = false;
    rc = helm_iteration(c, tol, max_iters, s);
    if (rc == EVO_OK) {
        helm::k_set_while_condition_helm<<<1, 32, 0, s>>>(handle, c->d_helm);
        c->launch_counter++;
    }
    e = cudaStreamEndCapture(s, nullptr);
    cap.open = false;
    if (rc != EVO_OK) return rc;
    if (e != cudaSuccess) return fail(EVO_ERR_CUDA, "helmholtz body capture: %s", cudaGetErrorString(e));
    c->kernels_per_cycle = c->launch_counter;
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, g, 0);
    if (e != cudaSuccess) return fail(EVO_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    c->helm_graph = gg.release(); c->helm_exec = exec; c->helm_tol = tol; c->helm_max_iters = max_iters;
    memcpy(&c->helm_A_graph, &c->helm_A, sizeof(OpSten));
    return EVO_OK;
}

extern "C" int evo_helmholtz_solve(evo_cycle *c, const evo_level_operator *A, const evo_solve_params *prm, evo_solve_result *res,
                                   double *res_hist)
{
    if (!c || !A || !prm || !res) return fail(EVO_ERR_INVALID, "null argument");
    if (c->p->desc.kind != EVO_PROBLEM_HELMHOLTZ || c->p->desc.scalar_words != 2) return fail(EVO_ERR_INVALID, "not a Helmholtz problem");
    if (prm->max_iters < 0 || prm->max_iters > 1000000) return fail(EVO_ERR_INVALID, "max_iters out of range");
    CU(cudaSetDevice(c->p->desc.device));
    memset(&c->helm_A, 0, sizeof(c->helm_A));
    {
        Sten &sA = c->helm_A.s[0][0];
        for (int q = 0; q < 27; ++q) {
            double re = A->coef[0][0][q][0], im = A->coef[0][0][q][1];
            if (re == 0.0 && im == 0.0) continue;
            if (q / 9 - 1 != 0) return fail(EVO_ERR_INVALID, "3-D stencil entry in a 2-D problem");
            int k = sA.nnz++;
            sA.ox[k] = (signed char)(q % 3 - 1); sA.oy[k] = (signed char)((q / 3) % 3 - 1); sA.oz[k] = 0;
            sA.re[k] = re; sA.im[k] = im;
        }
    }
    EV(ensure_hist(c, prm->max_iters));
    EV(build_helm_graph(c, prm->tol, prm->max_iters));
    const int samples = prm->samples > 0 ? prm->samples : 1;
    std::vector<float> times;
    memset(res, 0, sizeof(*res));
    helm::HelmState hs;
    for (int sidx = 0; sidx < samples; ++sidx) {
        cudaStream_t s = c->stream;
        EV(reset_cycle(c, s));
        if (prm->timeout_ms > 0) k_set_timeout_helm<<<1, 1, 0, s>>>(c->d_helm, (unsigned long long)prm->timeout_ms * 1000000ull);
        CU(cudaEventRecord(c->ev0, s));
        CU(cudaGraphLaunch(c->helm_exec, s));
        CU(cudaEventRecord(c->ev1, s));
        CU(cudaMemcpyAsync(&hs, c->d_helm, sizeof(hs), cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(c->h_hist, c->d_hist, sizeof(double) * (prm->max_iters + 1), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        times.push_back(ms);
    }
    res->status = hs.bad ? 1 : (hs.timed_out ? 2 : 0);
    res->iterations = hs.it;
    res->initial_residual = hs.init;
    res->final_residual = hs.cur;
    res->kernel_launches = c->kernels_prologue + c->kernels_per_cycle * (int64_t)hs.it;
    if (res_hist) memcpy(res_hist, c->h_hist, sizeof(double) * (hs.it + 1));
    std::sort(times.begin(), times.end());
    res->time_ms = times[times.size() / 2];
    res->time_ms_min = times[0];
    return EVO_OK;
}


// one solve of a cycle whose graph exists, alone on the device, stopped after `cap` iterations (0: to convergence)
static int solo_run(evo_cycle *c, int cap, float *ms)
{
    cudaStream_t s = c->stream;
    EV(reset_cycle(c, s));
    c->pristine = false;
    if (cap > 0) k_set_cap<<<1, 1, 0, s>>>(c->d_state, cap);
    CU(cudaEventRecord(c->ev0, s));
    CU(cudaGraphLaunch(c->exec, s));
    CU(cudaEventRecord(c->ev1, s));
    CU(cudaStreamSynchronize(s));
    CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return EVO_OK;
}

// The batch / pipeline times include the contention of everything else that was in flight.  Time is an optimisation
// objective (optimization/program.py:413, :449), so a member that converged is timed again with the GPU to itself: the
// whole solve when it is short, else 1 and 4 iterations (SolveState::cap) extrapolated to its iteration count.  Members
// that hit the iteration limit or diverged keep their time (unused: their fitness is the sentinel).
static int retime_solo(evo_cycle *c, const evo_solve_params *prm, evo_solve_result *res)
{
    const int its = res->iterations;
    if (res->status != 0 || its < 1 || its >= prm->max_iters) return EVO_OK;
    CU(cudaSetDevice(c->p->desc.device));
    float t_full = 0.f;
    if (its <= 4) {
        EV(solo_run(c, 0, &t_full));
        res->time_ms = res->time_ms_min = t_full;
    } else {
        // ONE run capped at 3 iterations: its span (CUDA events) + the remaining iterations at the per-iteration time the
        // device-side loop stamped (%globaltimer after iterations 1 and 3)
        const int cap = 3;
        EV(solo_run(c, cap, &t_full));
        CU(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(SolveState), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const SolveState &st = *c->h_state;
        double per_it = 0.0;
        if (st.it >= 2 && st.t_last > st.t_it1) per_it = (double)(st.t_last - st.t_it1) * 1e-6 / (st.it - 1);
        else per_it = t_full / cap;
        res->time_ms = res->time_ms_min = t_full + per_it * (its - st.it);
    }
    return EVO_OK;
}

extern "C" int evo_cycle_solve_begin(evo_cycle *c, const evo_solve_params *prm)
{
    if (!c || !prm) return fail(EVO_ERR_INVALID, "null argument");
    if (prm->flags & EVO_SOLVE_NO_GRAPH) return fail(EVO_ERR_UNSUPPORTED, "the solve pipeline needs the device-side outer loop");
    EV(prepare_solve(c, prm));
    return enqueue_solve(c, prm);
}

extern "C" int evo_cycle_solve_end(evo_cycle *c, const evo_solve_params *prm, evo_solve_result *res, double *res_hist)
{
    if (!c || !prm || !res) return fail(EVO_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->p->desc.device));
    memset(res, 0, sizeof(*res));
    float ms = 0.f;
    EV(collect_solve(c, prm, res, res_hist, &ms));
    res->time_ms = res->time_ms_min = ms;
    return EVO_OK;
}

extern "C" int evo_cycle_solve_retime(evo_cycle *c, const evo_solve_params *prm, evo_solve_result *res)
{
    if (!c || !prm || !res) return fail(EVO_ERR_INVALID, "null argument");
    if (!c->exec) return fail(EVO_ERR_INVALID, "no finished solve to time again");
    return retime_solo(c, prm, res);
}

// population evaluation on one GPU: every individual's complete solve is one graph launch on its own
// stream, nothing synchronises until all are in flight (reference: the sequential toolbox.map of
// optimization/program.py:491, one subprocess chain per individual)
extern "C" int evo_batch_solve(evo_cycle **cycles, int n, const evo_solve_params *prm, evo_solve_result *results,
                               double *res_hist, double *batch_ms)
{
    if (!cycles || !prm || !results || n < 0) return fail(EVO_ERR_INVALID, "null argument");
    if (n == 0) { if (batch_ms) *batch_ms = 0.0; return EVO_OK; }
    if (prm->flags & EVO_SOLVE_NO_GRAPH) return fail(EVO_ERR_UNSUPPORTED, "batch solve needs the device-side outer loop");
    const int samples = prm->samples > 0 ? prm->samples : 1;
    std::vector<std::vector<float>> times(n);
    std::vector<float> wall;
    cudaEvent_t b0, b1;
    CU(cudaEventCreate(&b0));
    CU(cudaEventCreate(&b1));
    cudaStream_t s0 = cycles[0]->stream;
    for (int sidx = 0; sidx < samples; ++sidx) {
        // fork: all streams wait for b0 on stream 0; join: stream 0 waits for every ev1
        CU(cudaEventRecord(b0, s0));
        for (int i = 0; i < n; ++i) {
            // graph capture / instantiation of individual i overlaps with the solves already in flight
            EV(prepare_solve(cycles[i], prm));
            if (i) CU(cudaStreamWaitEvent(cycles[i]->stream, b0, 0));
            EV(enqueue_solve(cycles[i], prm));
        }
        for (int i = 1; i < n; ++i) CU(cudaStreamWaitEvent(s0, cycles[i]->ev1, 0));
        CU(cudaEventRecord(b1, s0));
        for (int i = 0; i < n; ++i) {
            float ms = 0.f;
            EV(collect_solve(cycles[i], prm, &results[i], res_hist ? res_hist + (size_t)i * (prm->max_iters + 1) : nullptr, &ms));
            times[i].push_back(ms);
        }
        CU(cudaEventSynchronize(b1));
        float w = 0.f;
        CU(cudaEventElapsedTime(&w, b0, b1));
        wall.push_back(w);
    }
    for (int i = 0; i < n; ++i) {
        std::sort(times[i].begin(), times[i].end());
        results[i].time_ms = times[i][times[i].size() / 2];
        results[i].time_ms_min = times[i][0];
    }
    if (prm->flags & EVO_SOLVE_SOLO_TIMING)
        for (int i = 0; i < n; ++i) EV(retime_solo(cycles[i], prm, &results[i]));
    std::sort(wall.begin(), wall.end());
    if (batch_ms) *batch_ms = wall[wall.size() / 2];
    cudaEventDestroy(b0);
    cudaEventDestroy(b1);
    return EVO_OK;
}
