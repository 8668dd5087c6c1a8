// evo_dispatch.cuh -- one statement of the op list -> kernel launches (every function only enqueues work on `s`;
// safe under stream capture).  Included by evo_dispatch_inst.cu only.
#pragma once
#include "evo_runtime_internal.cuh"
#include "evo_kernels_run.cuh"

// grids up to this many inner nodes are latency bound: single-CTA / warp-per-node variants of the generic kernels
static constexpr long long SMALL_GRID_NODES = 4096;

template <typename T, int DIM, int NF> struct Launch {
    // d_partials holds the canonical row sums [NF][nzi][ni]; reduce them to SolveState::sum
    static int reduce_rows(evo_cycle *c, int ni, cudaStream_t s)
    {
        const double *vals = c->d_partials;
        if (DIM == 3) {
            double *planes = c->d_partials + (size_t)NF * ni * ni;
            k_reduce_planes<<<(unsigned)((NF * ni + 7) / 8), 256, 0, s>>>(c->d_partials, NF * ni, ni, planes);
            c->launch_counter += 1;
            vals = planes;
        }
        const evo_cycle::Finish &fin = c->fin;
        if (fin.on)     // solver graph: the loop's bookkeeping rides on the last reduction
            k_reduce_final_update<<<1, 32, 0, s>>>(vals, NF, ni, c->d_state, c->d_hist, fin.tol, fin.max_iters, fin.mode, fin.n_handles,
                                                   fin.h[0], fin.h[1]);
        else k_reduce_final<<<1, 32, 0, s>>>(vals, NF, ni, c->d_state);
        c->launch_counter += 1;
        CU(cudaGetLastError());
        return EVO_OK;
    }

    static int residual(evo_cycle *c, int l, bool norm, cudaStream_t s)
    {
        Geom g = c->p->geom[l];
        if (c->zc_lo >= 0 && !norm) { g.zlo = c->zc_lo; g.zhi = c->zc_hi; }   // domain decomposition: include ghost planes
        auto u = fields_of<T>(c->lv[l].buf[EVO_BUF_SOL], NF), f = fields_of<T>(c->lv[l].buf[EVO_BUF_RHS], NF),
             r = fields_of<T>(c->lv[l].buf[EVO_BUF_RES], NF);
        const int ni = g.n - 2;
        if (norm && star::try_residual_norm<T, DIM, NF>(g, c->sten[l], u, f, r, c->d_partials, !c->res_dead_on_entry, s)) {
            // residual and canonical row sums in one pass; the field itself is only stored if a later
            // statement may read it
            c->launch_counter += 1;
            EV(reduce_rows(c, ni, s));
            return EVO_OK;
        }
        if (norm && DIM == 2 && !slab_level(c->p, l)) {
            // 2-D (and complex) convergence norm: generic residual and row sums in one launch
            const unsigned nb = (unsigned)((ni + 3) / 4);
            if (c->res_dead_on_entry) k_residual_rowsum<T, DIM, NF, false><<<nb, 128, 0, s>>>(g, c->sten[l], u, f, r, c->d_partials);
            else k_residual_rowsum<T, DIM, NF, true><<<nb, 128, 0, s>>>(g, c->sten[l], u, f, r, c->d_partials);
            c->launch_counter += 1;
            EV(reduce_rows(c, ni, s));
            return EVO_OK;
        }
        if (!star::try_residual<T, DIM, NF>(c->p->sm_count, g, c->sten[l], u, f, r, s)) {
            if (slab_level(c->p, l)) return fail(EVO_ERR_UNSUPPORTED, "domain decomposition: residual needs the 7-point fast path");
            k_residual<T, DIM, NF><<<row_grid(g), BX, 0, s>>>(g, c->sten[l], u, f, r);
        }
        c->launch_counter++;
        if (norm) {
            if (slab_level(c->p, l)) return fail(EVO_ERR_UNSUPPORTED, "domain decomposition: use evo_cycle_residual_plane_sums");
            const long long nrows = (long long)ni * (DIM == 3 ? ni : 1);
            k_row_sumsq<T, DIM, NF><<<(unsigned)((nrows + 3) / 4), 128, 0, s>>>(g, r, c->d_partials);
            c->launch_counter += 1;
            EV(reduce_rows(c, ni, s));
        }
        CU(cudaGetLastError());
        return EVO_OK;
    }

    template <int NU> static int smooth_nu(evo_cycle *c, const evo_op &op, cudaStream_t s)
    {
        const int l = op.level;
        const Geom &g = c->p->geom[l];
        SmoothParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.nu = NU;
        sp.omega = op.omega;
        sp.write_all = 0;
        bool has_own[EVO_MAX_FIELDS] = {false, false}, written[EVO_MAX_FIELDS] = {false, false};
        for (int a = 0; a < NU; ++a) {
            sp.field[a] = op.unk_field[a];
            if (sp.field[a] < 0 || sp.field[a] >= NF) return fail(EVO_ERR_INVALID, "unknown refers to field %d", sp.field[a]);
            for (int d = 0; d < 3; ++d) sp.off[a][d] = d < DIM ? op.unk_off[a][d] : 0;
            written[sp.field[a]] = true;
            if (sp.off[a][0] == 0 && sp.off[a][1] == 0 && sp.off[a][2] == 0) has_own[sp.field[a]] = true;
        }
        for (int i = 0; i < NF; ++i)
            if (written[i] && !has_own[i])
                return fail(EVO_ERR_UNSUPPORTED, "local system without an unknown at the anchor node for field %d", i);
        auto rhs = fields_of<T>(c->lv[l].buf[EVO_BUF_RHS], NF);
        int reps = op.count > 1 ? op.count : 1;
        Geom gsub = g;   // domain decomposition: a sub-range of the owned planes (boundary planes first, interior later)
        if (c->zc_lo >= 0 && slab_level(c->p, l)) { gsub.zlo = c->zc_lo; gsub.zhi = c->zc_hi; }
        if constexpr (DIM == 2 && NF == 1 && NU == 1 && std::is_same<T, double>::value) {
            // large 2-D grids, 5-point star: register-streamed warp kernels, up to two sweeps per pass over HBM (temporal
            // blocking of the consecutive identical statements merged at build time); out of place into the [next] slot
            if ((op.mode == EVO_SMOOTH_JACOBI || op.mode == EVO_SMOOTH_REDBLACK) && c->lv[l].slot[0] && !slab_level(c->p, l) &&
                w2::star5_applicable(g, c->sten[l])) {
                while (reps > 0) {
                    const int k = reps >= 2 ? 2 : 1;
                    const double *src = (const double *)c->lv[l].buf[EVO_BUF_SOL][0], *fp = (const double *)rhs.p[0];
                    double *dst = (double *)c->lv[l].slot[0];
                    const bool ok = op.mode == EVO_SMOOTH_JACOBI ? w2::try_jacobi(c->p->sm_count, g, c->sten[l], src, fp, dst, op.omega, k, s)
                                                                 : w2::try_rbgs(c->p->sm_count, g, c->sten[l], src, fp, dst, op.omega, k, s);
                    if (!ok) return fail(EVO_ERR_CUDA, "2-D streaming sweep: launch failed");
                    c->launch_counter++;
                    const bool cor_alias = c->lv[l].buf[EVO_BUF_COR][0] == c->lv[l].buf[EVO_BUF_SOL][0];
                    std::swap(c->lv[l].buf[EVO_BUF_SOL][0], c->lv[l].slot[0]);
                    if (cor_alias) c->lv[l].buf[EVO_BUF_COR][0] = c->lv[l].buf[EVO_BUF_SOL][0];
                    c->lv[l].swapped[0] = !c->lv[l].swapped[0];
                    reps -= k;
                }
                CU(cudaGetLastError());
                return EVO_OK;
            }
        }
        if (NU == 1 && op.mode == EVO_SMOOTH_REDBLACK && c->lv[l].slot[0] && star::rbgs_stream_applicable<T, DIM, NF>(g, c->sten[l])) {
            // fused streaming kernel: up to 2 sweeps per pass, out of place into the [next] slot
            while (reps > 0) {
                // one launch per sweep: the 2-sweep variant (S = 4) is currently slower than two S = 2 launches
                const bool fuse2 = option(OPT_RB_FUSE2) != 0;
                const int k = (fuse2 && reps >= 2) ? 2 : 1;
                auto src = fields_of<T>(c->lv[l].buf[EVO_BUF_SOL], NF), dst = fields_of<T>(c->lv[l].slot, NF);
                if (!star::try_rbgs_stream<T, DIM, NF>(c->p->sm_count, gsub, c->sten[l], src, rhs, dst, op.omega, k, s)) break;
                c->launch_counter++;
                if (c->part_no_swap) { reps -= k; continue; }
                bool cor_alias = c->lv[l].buf[EVO_BUF_COR][0] == c->lv[l].buf[EVO_BUF_SOL][0];
                std::swap(c->lv[l].buf[EVO_BUF_SOL][0], c->lv[l].slot[0]);
                if (cor_alias) c->lv[l].buf[EVO_BUF_COR][0] = c->lv[l].buf[EVO_BUF_SOL][0];
                c->lv[l].swapped[0] = !c->lv[l].swapped[0];
                reps -= k;
            }
            if (reps == 0) { CU(cudaGetLastError()); return EVO_OK; }
        }
        if (slab_level(c->p, l) && !(NU == 1 && op.mode == EVO_SMOOTH_JACOBI))
            return fail(EVO_ERR_UNSUPPORTED, "domain decomposition: only pointwise RB-GS / Jacobi on 7-point stencils");
        for (int rep = 0; rep < reps; ++rep) {
            if (op.mode == EVO_SMOOTH_JACOBI) {
                // read the current slot, write the next slot, then `advance` (swap) the written fields
                void *cur[EVO_MAX_FIELDS], *nxt[EVO_MAX_FIELDS];
                for (int i = 0; i < NF; ++i) {
                    cur[i] = c->lv[l].buf[EVO_BUF_SOL][i];
                    nxt[i] = written[i] ? c->lv[l].slot[i] : cur[i];
                    if (written[i] && !nxt[i]) return fail(EVO_ERR_INVALID, "missing jacobi slot");
                }
                sp.color = -1;
                auto src = fields_of<T>(cur, NF), dst = fields_of<T>(nxt, NF);
                if (!(NU == 1 && star::try_smooth_point<T, DIM, NF>(c->p->sm_count, gsub, c->sten[l], sp, src, dst, rhs, s))) {
                    if (slab_level(c->p, l)) return fail(EVO_ERR_UNSUPPORTED, "domain decomposition: Jacobi needs the 7-point fast path");
                    k_smooth<T, DIM, NF, NU><<<row_grid(g), BX, 0, s>>>(g, c->sten[l], sp, src, dst, rhs);
                }
                c->launch_counter++;
                for (int i = 0; i < NF; ++i)
                    if (written[i] && !c->part_no_swap) {
                        bool cor_alias = c->lv[l].buf[EVO_BUF_COR][i] == c->lv[l].buf[EVO_BUF_SOL][i];
                        std::swap(c->lv[l].buf[EVO_BUF_SOL][i], c->lv[l].slot[i]);
                        c->lv[l].swapped[i] = !c->lv[l].swapped[i];
                        if (cor_alias) c->lv[l].buf[EVO_BUF_COR][i] = c->lv[l].buf[EVO_BUF_SOL][i];
                    }
            } else if (op.mode == EVO_SMOOTH_REDBLACK) {
                auto u = fields_of<T>(c->lv[l].buf[EVO_BUF_SOL], NF);
                if (color_order_dependent(c->sten[l], sp, NF)) {
                    if constexpr (NU == NF) {
                        bool done = false;
                        if constexpr (DIM == 2 && std::is_same<T, double>::value) {
                            // unknown a must be field a at the anchor (collective pointwise smoother)
                            bool canonical = true;
                            for (int a = 0; a < NU; ++a)
                                if (sp.field[a] != a || sp.off[a][0] || sp.off[a][1] || sp.off[a][2]) canonical = false;
                            const size_t smem = (size_t)5 * 2 * NF * g.pitch * sizeof(double);
                            const bool disabled = option(OPT_ROWSEQ_GLOBAL) != 0;
                            const bool no_pipe = option(OPT_ROWSEQ_NOPIPE) != 0;
                            // pipelined passes: as many of the remaining sweeps as the window fits (<= 4)
                            int k = std::min(reps - rep, 4);
                            while (k > 0 && (size_t)(4 * k + 3) * 2 * NF * g.pitch * sizeof(double) > 200 * 1024) --k;
                            Dense9<NF> dn;
                            bool dense_ok = true;
                            for (int a = 0; a < NF; ++a)
                                for (int j = 0; j < NF; ++j) {
                                    for (int q = 0; q < 9; ++q) dn.w[a][j][q] = 0.0;
                                    const Sten &sj = c->sten[l].s[a][j];
                                    for (int q = 0; q < sj.nnz; ++q) {
                                        if (sj.oz[q] != 0 || sj.im[q] != 0.0 || sj.re[q] == 0.0) dense_ok = false;
                                        dn.w[a][j][(sj.oy[q] + 1) * 3 + (sj.ox[q] + 1)] = sj.re[q];
                                    }
                                }
                            if (canonical && !disabled && !no_pipe && k > 0 && dense_ok) {
                                static bool attr_p = false;
                                if (!attr_p) {
                                    CU(cudaFuncSetAttribute(k2_smooth_rowseq_pipe<NF, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                                    CU(cudaFuncSetAttribute(k2_smooth_rowseq_pipe<NF, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                                    attr_p = true;
                                }
                                const size_t sm2 = (size_t)(4 * k + 3) * 2 * NF * g.pitch * sizeof(double);
                                if (dense9_pattern<NF>(dn) == 1)      // star / corner blocks: static sparsity (half the dependent adds)
                                    k2_smooth_rowseq_pipe<NF, 1><<<1, ROWSEQ_NT, sm2, s>>>(g, c->sten[l], dn, sp.omega, u, rhs, k);
                                else k2_smooth_rowseq_pipe<NF, 0><<<1, ROWSEQ_NT, sm2, s>>>(g, c->sten[l], dn, sp.omega, u, rhs, k);
                                done = true;
                                rep += k - 1;
                            } else if (canonical && !disabled && smem <= 200 * 1024) {
                                static bool attr = false;
                                if (!attr) {
                                    CU(cudaFuncSetAttribute(k2_smooth_rowseq_win<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                                    attr = true;
                                }
                                k2_smooth_rowseq_win<NF><<<1, 1024, smem, s>>>(g, c->sten[l], sp.omega, u, rhs);
                                done = true;
                            }
                        }
                        if (!done) k_smooth_rowseq<T, DIM, NF, NU><<<1, 1024, 0, s>>>(g, c->sten[l], sp, u, rhs);
                        c->launch_counter++;
                    } else {
                        return fail(EVO_ERR_UNSUPPORTED, "coloured block smoothers are not generated by the grammar");
                    }
                } else if ((long long)(g.n - 2) * (g.n - 2) * (DIM == 3 ? g.n - 2 : 1) <= SMALL_GRID_NODES && !slab_level(c->p, l)) {
                    // tiny grid: every remaining sweep and both colours in one launch (5-/7-point star, pointwise: on
                    // shared memory with straight-line arithmetic; else the generic local_solve)
                    const bool pointwise = NU == 1 && sp.field[0] == 0 && !sp.off[0][0] && !sp.off[0][1] && !sp.off[0][2];
                    if (!(pointwise && small::try_rb_small<T, DIM, NF>(g, c->sten[l], u, rhs, sp.omega, reps - rep, s)))
                        k_smooth_rb_small<T, DIM, NF, NU><<<1, 1024, 0, s>>>(g, c->sten[l], sp, u, rhs, reps - rep);
                    c->launch_counter++;
                    rep = reps;
                } else {
                    for (int color = 0; color < 2; ++color) {
                        sp.color = color;
                        if (!(NU == 1 && star::try_smooth_point<T, DIM, NF>(c->p->sm_count, g, c->sten[l], sp, u, u, rhs, s)))
                            k_smooth<T, DIM, NF, NU><<<row_grid(g, 2), BX, 0, s>>>(g, c->sten[l], sp, u, u, rhs);
                        c->launch_counter++;
                    }
                }
            } else {
                // lexicographic in-place sweeps (model-based mode of the reference): all remaining repetitions in one launch
                if (slab_level(c->p, l)) return fail(EVO_ERR_UNSUPPORTED, "domain decomposition: lexicographic sweeps are not distributed");
                int a = 0, b = 0;
                lex_skew(c->sten[l], sp, NF, &a, &b);
                sp.color = -1;
                sp.write_all = 1;
                auto u = fields_of<T>(c->lv[l].buf[EVO_BUF_SOL], NF);
                const int ni = g.n - 2;
                // anchors per hyperplane (upper bound) -> threads; one cluster of up to 8 CTAs
                const long long per_plane = DIM == 3 ? (long long)ni * ni / (1 + a) + ni : (long long)ni / (a > 0 ? a : 1) + 1;
                int threads = per_plane >= 1024 ? 1024 : (int)((per_plane + 31) / 32 * 32);
                int ctas = (int)std::min<long long>(8, (per_plane + 1023) / 1024);
                if (option(OPT_LEX_VARIANT) == 1) ctas = 1;
                while (ctas & (ctas - 1)) ++ctas;   // cluster sizes: powers of two
                cudaLaunchConfig_t cfg;
                memset(&cfg, 0, sizeof(cfg));
                cfg.gridDim = dim3(ctas);
                cfg.blockDim = dim3(threads);
                cfg.stream = s;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                const int sweeps = reps - rep;
                CU(cudaLaunchKernelEx(&cfg, k_smooth_lex<T, DIM, NF, NU>, g, c->sten[l], sp, u, rhs, a, b, sweeps));
                c->launch_counter++;
                rep = reps;
            }
        }
        CU(cudaGetLastError());
        return EVO_OK;
    }

    // skew of the hyperplanes of a lexicographic sweep: t = x + a*y + b*z must order every pair of conflicting anchors
    // (one writes what the other reads or writes) like the sequential loop does (x fastest, then y, then z)
    static void lex_skew(const OpSten &st, const SmoothParams &sp, int nf, int *a_out, int *b_out)
    {
        struct P { int f, x, y, z; };
        std::vector<P> W, R;
        for (int m = 0; m < sp.nu; ++m) {
            const P w{sp.field[m], sp.off[m][0], sp.off[m][1], sp.off[m][2]};
            W.push_back(w);
            R.push_back(w);
            for (int j = 0; j < nf; ++j) {
                const Sten &sj = st.s[sp.field[m]][j];
                for (int q = 0; q < sj.nnz; ++q) R.push_back(P{j, w.x + sj.ox[q], w.y + sj.oy[q], w.z + sj.oz[q]});
            }
        }
        auto cdiv = [](int p, int q) { return p >= 0 ? (p + q - 1) / q : -((-p) / q); };   // ceil(p / q), q > 0
        int a = 0, b = 0;
        for (int pass = 0; pass < 2; ++pass)     // pass 0: a from displacements with dz = 0; pass 1: b (needs a)
            for (const P &w : W)
                for (const P &r : R) {
                    if (w.f != r.f) continue;
                    for (int sgn = -1; sgn <= 1; sgn += 2) {
                        const int dx = sgn * (w.x - r.x), dy = sgn * (w.y - r.y), dz = sgn * (w.z - r.z);
                        if (pass == 0 && dz == 0 && dy > 0) a = std::max(a, cdiv(1 - dx, dy));
                        if (pass == 1 && dz > 0) b = std::max(b, cdiv(1 - dx - a * dy, dz));
                    }
                }
        *a_out = a; *b_out = b;
    }

    static bool color_order_dependent(const OpSten &st, const SmoothParams &sp, int nf)
    {
        for (int a = 0; a < sp.nu; ++a)
            for (int j = 0; j < nf; ++j) {
                const Sten &sj = st.s[sp.field[a]][j];
                for (int q = 0; q < sj.nnz; ++q) {
                    int ox = sp.off[a][0] + sj.ox[q], oy = sp.off[a][1] + sj.oy[q], oz = sp.off[a][2] + sj.oz[q];
                    bool is_unknown = false;
                    for (int m = 0; m < sp.nu; ++m)
                        if (sp.field[m] == j && sp.off[m][0] == ox && sp.off[m][1] == oy && sp.off[m][2] == oz) is_unknown = true;
                    if (is_unknown) continue;
                    for (int m = 0; m < sp.nu; ++m) {
                        if (sp.field[m] != j) continue;
                        int sx = ox - sp.off[m][0], sy = oy - sp.off[m][1], sz = oz - sp.off[m][2];
                        if (((sx + sy + sz) & 1) == 0) return true;
                    }
                }
            }
        return false;
    }

    static int smooth(evo_cycle *c, const evo_op &op, cudaStream_t s)
    {
        switch (op.n_unknowns) {
        case 1: return smooth_nu<1>(c, op, s);
        case 2: return smooth_nu<2>(c, op, s);
        case 3: return smooth_nu<3>(c, op, s);
        case 4: return smooth_nu<4>(c, op, s);
        case 5: return smooth_nu<5>(c, op, s);
        case 6: return smooth_nu<6>(c, op, s);
        case 7: return smooth_nu<7>(c, op, s);
        case 8: return smooth_nu<8>(c, op, s);
        default: return fail(EVO_ERR_INVALID, "local system size %d", op.n_unknowns);
        }
    }

    static int restrict_(evo_cycle *c, const evo_op &op, cudaStream_t s)
    {
        const int l = op.level;
        const Geom &gf = c->p->geom[l];
        Geom gc = c->p->geom[l - 1];
        if (c->zc_lo >= 0) { gc.zlo = c->zc_lo; gc.zhi = c->zc_hi; }   // domain decomposition: only these coarse planes
        auto src = fields_of<T>(c->lv[l].buf[op.src], NF), dst = fields_of<T>(c->lv[l - 1].buf[op.dst], NF);
        k_restrict<T, DIM, NF><<<row_grid(gc), BX, 0, s>>>(gf, gc, c->p->R, src, dst);
        c->launch_counter++;
        CU(cudaGetLastError());
        return EVO_OK;
    }

    static int residual_restrict(evo_cycle *c, const evo_op &op, cudaStream_t s)
    {
        const int l = op.level;
        const Geom &gf = c->p->geom[l];
        Geom gc = c->p->geom[l - 1];
        if (c->zc_lo >= 0) { gc.zlo = c->zc_lo; gc.zhi = c->zc_hi; }   // domain decomposition: only these coarse planes
        auto u = fields_of<T>(c->lv[l].buf[EVO_BUF_SOL], NF), f = fields_of<T>(c->lv[l].buf[EVO_BUF_RHS], NF),
             dst = fields_of<T>(c->lv[l - 1].buf[EVO_BUF_RHS], NF);
        // `SOL@(l-1) = 0` as the next statement: the 2-D kernels store the zeros of the inner nodes along with the
        // coarse right-hand side (the boundary layer of a correction level is never written, it stays 0)
        Fields<T> zero;
        for (int i = 0; i < EVO_MAX_FIELDS; ++i) zero.p[i] = nullptr;
        if (c->fuse_zero)
            for (int i = 0; i < NF; ++i) zero.p[i] = (T *)c->lv[l - 1].buf[EVO_BUF_SOL][i];
        bool done2d = false;
        if constexpr (DIM == 2 && NF == 1 && std::is_same<T, double>::value)
            done2d = !slab_level(c->p, l) && w2::try_residual_restrict(c->p->sm_count, gf, gc, c->sten[l], c->p->R, (const double *)u.p[0],
                                                                      (const double *)f.p[0], (double *)dst.p[0], (double *)zero.p[0], s);
        if (done2d) {
            if (zero.p[0]) c->fuse_zero = false;
        } else if (!star::try_residual_restrict_col<T, DIM, NF>(c->p->sm_count, gf, gc, c->sten[l], c->p->R, u, f, dst, s) &&
                   !star::try_residual_restrict<T, DIM, NF>(c->p->sm_count, gf, gc, c->sten[l], c->p->R, u, f, dst, s)) {
            if (slab_level(c->p, l)) return fail(EVO_ERR_UNSUPPORTED, "domain decomposition: fused residual+restriction needs the fast path");
            const long long cn = (long long)(gc.n - 2) * (gc.n - 2) * (DIM == 3 ? gc.n - 2 : 1);
            // measured (scripts/op_costs.py): two fields, 63^2 coarse nodes: 6.9 us (warp) vs 13.4 us (thread per node); 127^2: 18.8 vs
            // 13.9 us; one field: the thread-per-node kernel wins from 63^2 on
            const long long warp_max = option(OPT_RR_WARP_NODES) > 0 ? option(OPT_RR_WARP_NODES)
                                                                      : (DIM == 3 ? SMALL_GRID_NODES / 8 : (NF >= 2 ? SMALL_GRID_NODES : SMALL_GRID_NODES / 4));
            if (cn <= warp_max)      // tiny coarse grid: one warp per coarse node
                k_residual_restrict_warp<T, DIM, NF><<<(unsigned)((cn + 3) / 4), 128, 0, s>>>(gf, gc, c->sten[l], c->p->R, u, f, dst, zero);
            else k_residual_restrict<T, DIM, NF><<<row_grid(gc), BX, 0, s>>>(gf, gc, c->sten[l], c->p->R, u, f, dst, zero);
            if (zero.p[0]) c->fuse_zero = false;
        }
        c->launch_counter++;
        CU(cudaGetLastError());
        return EVO_OK;
    }

    static int prolong(evo_cycle *c, const evo_op &op, bool add, cudaStream_t s)
    {
        const int l = op.level;
        Geom gf = c->p->geom[l];
        const Geom &gc = c->p->geom[l - 1];
        if (c->zc_lo >= 0) { gf.zlo = c->zc_lo; gf.zhi = c->zc_hi; }   // domain decomposition: include ghost planes
        auto src = fields_of<T>(c->lv[l - 1].buf[op.src], NF);
        auto dst = fields_of<T>(c->lv[l].buf[add ? EVO_BUF_SOL : op.dst], NF);
        if (add) {
            bool done2d = false;
            if constexpr (DIM == 2 && NF == 1 && std::is_same<T, double>::value)
                done2d = !slab_level(c->p, l) && w2::try_prolong_add(gf, gc, c->p->P, (const double *)src.p[0], (double *)dst.p[0], op.omega, s);
            if (!done2d && !star::try_prolong_add<T, DIM, NF>(c->p->sm_count, gf, gc, c->p->P, src, dst, op.omega, s)) {
                if (slab_level(c->p, l)) return fail(EVO_ERR_UNSUPPORTED, "domain decomposition: prolongation needs the fast path");
                k_prolong<T, DIM, NF, true><<<row_grid(gf), BX, 0, s>>>(gf, gc, c->p->P, src, dst, op.omega);
            }
        } else {
            if (slab_level(c->p, l)) return fail(EVO_ERR_UNSUPPORTED, "domain decomposition: PROLONG_SET not supported");
            k_prolong<T, DIM, NF, false><<<row_grid(gf), BX, 0, s>>>(gf, gc, c->p->P, src, dst, 1.0);
        }
        c->launch_counter++;
        CU(cudaGetLastError());
        return EVO_OK;
    }

    static int richardson(evo_cycle *c, const evo_op &op, cudaStream_t s)
    {
        if (slab_level(c->p, op.level)) return fail(EVO_ERR_UNSUPPORTED, "domain decomposition: Richardson steps not supported");
        // field by field: tmp = RHS_i - (A SOL)_i from the current values, then SOL_i += w * tmp
        const int l = op.level;
        const Geom &g = c->p->geom[l];
        auto u = fields_of<T>(c->lv[l].buf[EVO_BUF_SOL], NF), f = fields_of<T>(c->lv[l].buf[EVO_BUF_RHS], NF);
        for (int i = 0; i < NF; ++i) {
            T *tmp = (T *)c->lv[l].slot[i];
            if (!tmp) return fail(EVO_ERR_INVALID, "missing scratch slot");
            Fields<T> r = u;  // residual of field i goes to tmp; other entries unused (NF passes write all -> use scratch trick)
            for (int j = 0; j < NF; ++j) r.p[j] = (T *)c->lv[l].slot[j];
            k_residual<T, DIM, NF><<<row_grid(g), BX, 0, s>>>(g, c->sten[l], u, f, r);
            k_axpy_inner<T, DIM><<<row_grid(g), BX, 0, s>>>(g, (T *)c->lv[l].buf[EVO_BUF_SOL][i], tmp, op.omega);
            c->launch_counter += 2;
        }
        // the slot was used as scratch: restore its boundary invariant (inner values are don't-care)
        CU(cudaGetLastError());
        return EVO_OK;
    }

};

template <int DIM, int NF> static int coarse_cg(evo_cycle *c, const evo_op &op, cudaStream_t s)
{
    const int l = op.level;
    const Geom &g = c->p->geom[l];
    if (l != c->p->desc.min_level) return fail(EVO_ERR_UNSUPPORTED, "coarse-grid solver below its level");
    auto x = fields_of<double>(c->lv[l].buf[EVO_BUF_SOL], NF), b = fields_of<double>(c->lv[l].buf[EVO_BUF_RHS], NF);
    {
        // shared-memory resident variant when the four CG vectors fit
        const int ni = g.n - 2;
        const size_t vol = (size_t)g.n * g.n * (DIM == 3 ? g.n : 1);
        const size_t smem = 4 * NF * vol * sizeof(double);
        const int nrows = ni * (DIM == 3 ? ni : 1);
        const bool disabled = option(OPT_CG_GLOBAL) != 0;
        const bool no_reg = option(OPT_CG_NOREG) != 0;
        if constexpr (DIM == 2) {
            // one node per thread, CG vectors in registers (coarsest grids up to 33 x 33)
            Dense9<NF> dn;
            bool dense_ok = !disabled && !no_reg && ni <= 32;
            for (int a = 0; a < NF && dense_ok; ++a)
                for (int j = 0; j < NF; ++j) {
                    for (int q = 0; q < 9; ++q) dn.w[a][j][q] = 0.0;
                    const Sten &sj = c->sten[l].s[a][j];
                    for (int q = 0; q < sj.nnz; ++q) {
                        if (sj.oz[q] != 0 || sj.im[q] != 0.0 || sj.re[q] == 0.0) dense_ok = false;
                        dn.w[a][j][(sj.oy[q] + 1) * 3 + (sj.ox[q] + 1)] = sj.re[q];
                    }
                }
            if (dense_ok) {
                if (dense9_pattern<NF>(dn) == 1)
                    k2_coarse_cg_reg<NF, 1><<<1, 1024, NF * vol * sizeof(double), s>>>(g, dn, x, b, op.count, op.tol, c->d_cg_iters);
                else k2_coarse_cg_reg<NF, 0><<<1, 1024, NF * vol * sizeof(double), s>>>(g, dn, x, b, op.count, op.tol, c->d_cg_iters);
                c->launch_counter++;
                CU(cudaGetLastError());
                return EVO_OK;
            }
        }
        if (!disabled && smem <= 160 * 1024 && NF * nrows <= 512 && (DIM == 2 || ni <= 64)) {
            static bool attr = false;
            if (!attr) {
                CU(cudaFuncSetAttribute(k_coarse_cg_smem<DIM, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                attr = true;
            }
            k_coarse_cg_smem<DIM, NF><<<1, 1024, smem, s>>>(g, c->sten[l], x, b, op.count, op.tol, c->d_cg_iters);
            c->launch_counter++;
            CU(cudaGetLastError());
            return EVO_OK;
        }
    }
    auto r = fields_of<double>(c->krylov[0], NF), p = fields_of<double>(c->krylov[1], NF), ap = fields_of<double>(c->krylov[2], NF);
    k_coarse_cg<DIM, NF><<<1, 1024, 0, s>>>(g, c->sten[l], x, b, r, p, ap, (double *)c->krylov[3][0], op.count, op.tol,
                                            c->d_cg_iters);
    c->launch_counter++;
    CU(cudaGetLastError());
    return EVO_OK;
}

template <typename T, int DIM, int NF> int enqueue_op(evo_cycle *c, const evo_op &op, cudaStream_t s)
{
    using L = Launch<T, DIM, NF>;
    evo_problem *p = c->p;
    const int l = op.level;
    const size_t esz = sizeof(double) * p->words;
    switch (op.code) {
    case EVO_OP_ZERO:
        for (int i = 0; i < NF; ++i) CU(cudaMemsetAsync(c->lv[l].buf[op.dst][i], 0, (size_t)p->geom[l].total * esz, s));
        return EVO_OK;
    case EVO_OP_COPY:
        for (int i = 0; i < NF; ++i)
            if (c->lv[l].buf[op.dst][i] != c->lv[l].buf[op.src][i])
                CU(cudaMemcpyAsync(c->lv[l].buf[op.dst][i], c->lv[l].buf[op.src][i], (size_t)p->geom[l].total * esz,
                                   cudaMemcpyDeviceToDevice, s));
        return EVO_OK;
    case EVO_OP_RESIDUAL: return L::residual(c, l, false, s);
    case EVO_OP_RICHARDSON: return L::richardson(c, op, s);
    case EVO_OP_SMOOTH: return L::smooth(c, op, s);
    case EVO_OP_RESTRICT: return L::restrict_(c, op, s);
    case EVO_OP_RESIDUAL_RESTRICT: return L::residual_restrict(c, op, s);
    case EVO_OP_PROLONG_ADD: return L::prolong(c, op, true, s);
    case EVO_OP_PROLONG_SET: return L::prolong(c, op, false, s);
    case EVO_OP_COARSE_SOLVE:
        if constexpr (std::is_same<T, double>::value) {
            return coarse_cg<DIM, NF>(c, op, s);
        } else {
            if (l != p->desc.min_level) return fail(EVO_ERR_UNSUPPORTED, "coarse-grid solver below its level");
            const Geom &g = p->geom[l];
            helm::k2_coarse_bicgstab<<<1, 1024, 0, s>>>(
                g, c->sten[l], helm_rden(p, l), (cplx *)c->lv[l].buf[EVO_BUF_SOL][0], (const cplx *)c->lv[l].buf[EVO_BUF_RHS][0],
                (cplx *)c->lv[l].buf[EVO_BUF_RES][0], (cplx *)c->krylov[0][0], (cplx *)c->krylov[1][0], (cplx *)c->krylov[2][0],
                (cplx *)c->krylov[3][0], (cplx *)c->krylov[4][0], (cplx *)c->krylov[5][0], (cplx *)c->krylov[6][0], op.count, op.tol,
                p->desc.kind == EVO_PROBLEM_HELMHOLTZ ? 1 : 0);
            c->launch_counter++;
            CU(cudaGetLastError());
            return EVO_OK;
        }
    default: return fail(EVO_ERR_UNSUPPORTED, "op code %d not implemented", op.code);
    }
}


// ---- fused runs of statements on small levels (evo_kernels_run.cuh) -------------------------------------------------
constexpr size_t RUN_SMEM_BUDGET = 200 * 1024;
template <int DIM> inline size_t run_compact_doubles(const Geom &g) { return ((size_t)g.n * g.n * (DIM == 3 ? g.n : 1) + 1) & ~(size_t)1; }
// highest level of a shared-memory resident run (EVO_COARSE_FUSE = 5)
template <int DIM, int NF> int run_resident_cap(const evo_problem *p)
{
    const int lo = p->desc.min_level;
    size_t bytes = 4 * NF * run_compact_doubles<DIM>(p->geom[lo]) * sizeof(double);      // CG vectors
    int cap = lo - 1;
    for (int l = lo; l <= p->desc.max_level && l - lo < RUN_MAX_LEVELS; ++l) {
        bytes += 3 * NF * run_compact_doubles<DIM>(p->geom[l]) * sizeof(double);
        const long long ni = p->geom[l].n - 2;
        if (bytes > RUN_SMEM_BUDGET || ni * ni * (DIM == 3 ? ni : 1) > 16 * 1024) break;
        cap = l;
    }
    return cap;
}

// can this statement be part of a fused run?
template <typename T, int DIM, int NF> bool run_eligible(const evo_cycle *c, const evo_op &op)
{
    if constexpr (!std::is_same<T, double>::value) {
        return false;
    } else {
        const evo_problem *p = c->p;
        if (p->desc.kind != EVO_PROBLEM_LINEAR || p->slab_lc > 0 || option(OPT_COARSE_FUSE) == 0) return false;
        const int l = op.level;
        if (l < p->desc.min_level || l > p->desc.max_level) return false;
        if (l - p->desc.min_level >= RUN_MAX_LEVELS) return false;
        // EVO_COARSE_FUSE: 1 (default) 2-D up to 129^2 / 3-D up to 17^3 in one CTA; 2 also 257^2 / 33^3 (thread-block
        // cluster); 3 only up to 65^2 / 17^3; 4 only 3-D up to 17^3
        const int mode = option(OPT_COARSE_FUSE);
        const int nmax = DIM == 2 ? (mode == 2 ? 257 : (mode == 3 ? 65 : (mode == 4 ? 0 : 129))) : (mode == 2 ? 33 : 17);
        if (mode == 5) {
            // shared-memory resident runs: the levels from the coarsest one up whose arrays (three per field: SOL, its
            // [next] slot, RHS) fit into one CTA's shared memory together with the CG vectors
            if (l > run_resident_cap<DIM, NF>(p)) return false;
        } else if (p->geom[l].n > nmax) return false;
        switch (op.code) {
        case EVO_OP_ZERO: case EVO_OP_COPY: return true;
        case EVO_OP_RESIDUAL: return c->has_sten[l];
        case EVO_OP_RESTRICT: case EVO_OP_PROLONG_ADD: case EVO_OP_PROLONG_SET: return l > p->desc.min_level;
        case EVO_OP_RESIDUAL_RESTRICT: return l > p->desc.min_level && c->has_sten[l];
        case EVO_OP_COARSE_SOLVE: {
            // the shared-memory CG as a device function; one CTA must hold the level (<= 16 nodes per thread)
            const int ni = p->geom[l].n - 2, nrows = ni * (DIM == 3 ? ni : 1);
            const size_t vol = (size_t)p->geom[l].n * p->geom[l].n * (DIM == 3 ? p->geom[l].n : 1);
            return l == p->desc.min_level && c->has_sten[l] && option(OPT_CG_GLOBAL) == 0 && 4 * NF * vol * sizeof(double) <= 40 * 1024 &&
                   NF * nrows <= 512 && (long long)nrows * ni <= 16 * 1024;
        }
        case EVO_OP_SMOOTH: {
            if (op.kind != EVO_KIND_LINEAR || !c->has_sten[l] || op.n_unknowns < 1 || op.n_unknowns > EVO_MAX_UNKNOWNS) return false;
            SmoothParams sp;
            memset(&sp, 0, sizeof(sp));
            sp.nu = op.n_unknowns;
            bool has_own[EVO_MAX_FIELDS] = {false, false}, written[EVO_MAX_FIELDS] = {false, false};
            for (int a = 0; a < sp.nu; ++a) {
                sp.field[a] = op.unk_field[a];
                if (sp.field[a] < 0 || sp.field[a] >= NF) return false;
                for (int d = 0; d < 3; ++d) sp.off[a][d] = d < DIM ? op.unk_off[a][d] : 0;
                written[sp.field[a]] = true;
                if (sp.off[a][0] == 0 && sp.off[a][1] == 0 && sp.off[a][2] == 0) has_own[sp.field[a]] = true;
            }
            for (int i = 0; i < NF; ++i)
                if (written[i] && !has_own[i]) return false;
            if (op.mode == EVO_SMOOTH_JACOBI) {
                for (int i = 0; i < NF; ++i)
                    if (written[i] && !c->lv[l].slot[i]) return false;
                return true;
            }
            if (op.mode == EVO_SMOOTH_REDBLACK) return !Launch<T, DIM, NF>::color_order_dependent(c->sten[l], sp, NF);
            return false;
        }
        default: return false;
        }
    }
}

// enqueue ops[0..n) (all run_eligible) as fused launches; buffer exchanges of out-of-place statements are applied to
// the cycle's pointers exactly as the stand-alone dispatch applies them
template <typename T, int DIM, int NF> int enqueue_run(evo_cycle *c, const evo_op *ops, int n, cudaStream_t s)
{
    if constexpr (!std::is_same<T, double>::value) {
        return fail(EVO_ERR_UNSUPPORTED, "fused runs: real problems only");
    } else {
        evo_problem *p = c->p;
        const int lbase = p->desc.min_level;
        for (int i0 = 0; i0 < n; i0 += RUN_MAX_OPS) {
            const int m = std::min(RUN_MAX_OPS, n - i0);
            RunTable tab;
            static_assert(sizeof(RunTable) <= 4000, "the run table travels as a kernel parameter");
            memset(&tab, 0, sizeof(tab));
            tab.n = m;
            tab.lbase = lbase;
            tab.sten = c->d_run_sten;
            tab.sp = c->d_run_sp;
            tab.rp = c->d_run_rp;
            long long nodes_max = 1;
            size_t cg_smem = 0;      // > 0: the run holds a coarse-grid solve (one CTA, dynamic shared memory for the CG vectors)
            for (int q = 0; q < m; ++q) {
                const evo_op &op = ops[i0 + q];
                const int l = op.level;
                RunOp &r = tab.op[q];
                r.code = op.code;
                r.li = l - lbase;
                r.lj = l - 1 - lbase;
                r.reps = op.count > 1 ? op.count : 1;
                r.mode = op.mode;
                r.omega = op.omega;
                tab.geom[r.li] = p->geom[l];
                if (r.lj >= 0) tab.geom[r.lj] = p->geom[l - 1];
                const Geom &g = p->geom[l];
                nodes_max = std::max(nodes_max, (long long)(g.n - 2) * (g.n - 2) * (DIM == 3 ? g.n - 2 : 1));
                LevelMem &lv = c->lv[l];
                switch (op.code) {
                case EVO_OP_ZERO:
                    for (int f = 0; f < NF; ++f) r.b[f] = lv.buf[op.dst][f];
                    break;
                case EVO_OP_COPY:
                    for (int f = 0; f < NF; ++f) { r.a[f] = lv.buf[op.src][f]; r.b[f] = lv.buf[op.dst][f]; }
                    break;
                case EVO_OP_RESIDUAL:
                    for (int f = 0; f < NF; ++f) { r.a[f] = lv.buf[EVO_BUF_SOL][f]; r.b[f] = lv.buf[EVO_BUF_RHS][f]; r.c[f] = lv.buf[EVO_BUF_RES][f]; }
                    break;
                case EVO_OP_RESTRICT:
                    for (int f = 0; f < NF; ++f) { r.a[f] = lv.buf[op.src][f]; r.b[f] = c->lv[l - 1].buf[op.dst][f]; }
                    break;
                case EVO_OP_RESIDUAL_RESTRICT:
                    for (int f = 0; f < NF; ++f) {
                        r.a[f] = lv.buf[EVO_BUF_SOL][f]; r.c[f] = lv.buf[EVO_BUF_RHS][f]; r.b[f] = c->lv[l - 1].buf[EVO_BUF_RHS][f];
                    }
                    break;
                case EVO_OP_PROLONG_ADD:
                case EVO_OP_PROLONG_SET:
                    for (int f = 0; f < NF; ++f) {
                        r.a[f] = c->lv[l - 1].buf[op.src][f];
                        r.b[f] = lv.buf[op.code == EVO_OP_PROLONG_ADD ? EVO_BUF_SOL : op.dst][f];
                    }
                    break;
                case EVO_OP_COARSE_SOLVE:
                    for (int f = 0; f < NF; ++f) { r.a[f] = lv.buf[EVO_BUF_SOL][f]; r.b[f] = lv.buf[EVO_BUF_RHS][f]; }
                    r.c[0] = c->d_cg_iters;
                    r.reps = op.count;
                    r.omega = op.tol;
                    cg_smem = std::max(cg_smem, 4 * NF * (size_t)g.n * g.n * (DIM == 3 ? g.n : 1) * sizeof(double));
                    break;
                case EVO_OP_SMOOTH: {
                    const ptrdiff_t idx = &op - c->ops.data();
                    if (idx < 0 || idx >= (ptrdiff_t)c->ops.size()) return fail(EVO_ERR_INVALID, "fused runs take statements of the cycle");
                    r.spi = (int)idx;
                    r.nu = op.n_unknowns;
                    for (int a = 0; a < op.n_unknowns; ++a) r.written |= 1u << op.unk_field[a];
                    for (int f = 0; f < NF; ++f) { r.a[f] = lv.buf[EVO_BUF_SOL][f]; r.b[f] = lv.slot[f]; r.c[f] = lv.buf[EVO_BUF_RHS][f]; }
                    if (op.mode == EVO_SMOOTH_JACOBI && (r.reps & 1)) {
                        // an odd number of out-of-place sweeps leaves the result in the [next] slots
                        for (int f = 0; f < NF; ++f)
                            if ((r.written >> f) & 1u) {
                                const bool cor_alias = lv.buf[EVO_BUF_COR][f] == lv.buf[EVO_BUF_SOL][f];
                                std::swap(lv.buf[EVO_BUF_SOL][f], lv.slot[f]);
                                lv.swapped[f] = !lv.swapped[f];
                                if (cor_alias) lv.buf[EVO_BUF_COR][f] = lv.buf[EVO_BUF_SOL][f];
                            }
                    }
                    break;
                }
                default: return fail(EVO_ERR_INVALID, "statement %d cannot be part of a fused run", op.code);
                }
            }
            int numax = 1;
            for (int q = 0; q < m; ++q)
                if (tab.op[q].code == EVO_OP_SMOOTH) numax = std::max(numax, tab.op[q].nu);
            // EVO_COARSE_FUSE = 5: every field array the run touches lives in shared memory for the whole launch
            size_t run_smem = cg_smem;
            if (option(OPT_COARSE_FUSE) == 5) {
                int n_stage = 0, off = 0;
                bool fits = true;
                auto stage = [&](void *&ptr, int li) {        // replace a device pointer by its shared-memory tag
                    if (!ptr || !fits) return;
                    for (int e = 0; e < n_stage; ++e)
                        if (tab.stage[e].ptr == ptr && tab.stage[e].li == li) { ptr = run_tag(tab.stage[e].off); return; }
                    if (n_stage == RUN_MAX_STAGE) { fits = false; return; }
                    tab.stage[n_stage] = RunStage{ptr, li, off};
                    ptr = run_tag(off);
                    off += (int)run_compact_doubles<DIM>(tab.geom[li]);
                    ++n_stage;
                };
                RunTable saved = tab;
                for (int q = 0; q < m && fits; ++q) {
                    RunOp &r = tab.op[q];
                    for (int f = 0; f < NF; ++f) {
                        switch (r.code) {
                        case EVO_OP_ZERO: stage(r.b[f], r.li); break;
                        case EVO_OP_COPY: stage(r.a[f], r.li); stage(r.b[f], r.li); break;
                        case EVO_OP_RESIDUAL: case EVO_OP_SMOOTH: stage(r.a[f], r.li); stage(r.b[f], r.li); stage(r.c[f], r.li); break;
                        case EVO_OP_RESTRICT: stage(r.a[f], r.li); stage(r.b[f], r.lj); break;
                        case EVO_OP_RESIDUAL_RESTRICT: stage(r.a[f], r.li); stage(r.c[f], r.li); stage(r.b[f], r.lj); break;
                        case EVO_OP_PROLONG_ADD: case EVO_OP_PROLONG_SET: stage(r.a[f], r.lj); stage(r.b[f], r.li); break;
                        case EVO_OP_COARSE_SOLVE: stage(r.a[f], r.li); stage(r.b[f], r.li); break;
                        default: fits = false; break;
                        }
                    }
                }
                const size_t bytes = (size_t)off * sizeof(double) + cg_smem;
                if (fits && bytes <= RUN_SMEM_BUDGET && nodes_max <= 16 * 1024) {
                    tab.n_stage = n_stage;
                    tab.cg_off = off;
                    run_smem = bytes;
                    for (int li = 0; li < RUN_MAX_LEVELS; ++li) {          // compact geometry: pitch = n
                        Geom &g = tab.geom[li];
                        if (g.n == 0) continue;
                        g.pitch = g.n;
                        g.plane = (long long)g.n * g.n;
                        g.total = g.plane * (DIM == 3 ? g.n : 1);
                    }
                } else {
                    tab = saved;       // does not fit: the run works on device memory as before
                }
            }
            // one CTA (block barriers) up to 16 nodes per thread, else a cluster of up to 8 CTAs
            const int tmax = numax <= 2 ? 1024 : 512;
            int ctas = 1;
            while (cg_smem == 0 && tab.n_stage == 0 && ctas < 8 && nodes_max > (long long)ctas * tmax * 16) ctas *= 2;
            const int threads = nodes_max >= tmax ? tmax : (int)((nodes_max + 31) / 32 * 32);
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(ctas);
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = run_smem;
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = ctas > 1 ? 1 : 0;
            if (ctas > 1) {
                if (numax <= 2) CU(cudaLaunchKernelEx(&cfg, k_run<DIM, NF, 2, true>, tab));
                else if (numax <= 4) CU(cudaLaunchKernelEx(&cfg, k_run<DIM, NF, 4, true>, tab));
                else CU(cudaLaunchKernelEx(&cfg, k_run<DIM, NF, 8, true>, tab));
            } else {
                static bool attr = false;
                if (!attr) {
                    CU(cudaFuncSetAttribute(k_run<DIM, NF, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RUN_SMEM_BUDGET));
                    CU(cudaFuncSetAttribute(k_run<DIM, NF, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RUN_SMEM_BUDGET));
                    CU(cudaFuncSetAttribute(k_run<DIM, NF, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RUN_SMEM_BUDGET));
                    attr = true;
                }
                if (numax <= 2) CU(cudaLaunchKernelEx(&cfg, k_run<DIM, NF, 2, false>, tab));
                else if (numax <= 4) CU(cudaLaunchKernelEx(&cfg, k_run<DIM, NF, 4, false>, tab));
                else CU(cudaLaunchKernelEx(&cfg, k_run<DIM, NF, 8, false>, tab));
            }
            c->launch_counter++;
        }
        return EVO_OK;
    }
}

template <typename T, int DIM, int NF> int op_residual(evo_cycle *c, int level, bool norm, cudaStream_t s)
{
    return Launch<T, DIM, NF>::residual(c, level, norm, s);
}
template <typename T, int DIM, int NF> int op_restrict(evo_cycle *c, const evo_op &op, cudaStream_t s)
{
    return Launch<T, DIM, NF>::restrict_(c, op, s);
}
template <typename T, int DIM, int NF> int op_reduce_rows(evo_cycle *c, int ni, cudaStream_t s)
{
    return Launch<T, DIM, NF>::reduce_rows(c, ni, s);
}
