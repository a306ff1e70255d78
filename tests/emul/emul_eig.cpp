// CPU replay of the per-thread solver bodies in bspatom_b200/csrc/bsp_core.h and
// of the stage schedule in bsp_driver.h.  TEST INFRASTRUCTURE ONLY: it exists so
// that `pytest -m "not gpu"` can check the solver logic (indexing, bracket
// bookkeeping, iteration schedule) without a GPU.  It is never built into
// libbspatom.so and nothing in the product imports it.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../bspatom_b200/csrc/bsp_core.h"
#include "../../bspatom_b200/csrc/bsp_driver.h"

template <int B>
struct EmulExec {
    BspEigChunk g;
    std::vector<double> cand_s;
    std::vector<int> cand_c;
    int cur_iter = 0, open_ok = 0;
    int ckpt = 0;   /* 1: the full-width solves (iterations 0, 1) run check-pointed, like the kernels for B <= 6 */
    void bounds() {
        cand_s.assign(g.npencil * BSP_NCAND, 0.0);
        cand_c.assign(g.npencil * BSP_NCAND, 0);
        for (int p = 0; p < g.npencil; ++p)
            for (int l = 0; l < BSP_NCAND; ++l) bsp_bounds_candidate<B>(g, p, l, cand_s.data(), cand_c.data());
        for (int p = 0; p < g.npencil; ++p) bsp_bounds_pick(g, p, cand_s.data(), cand_c.data());
    }
    void round(int r, int max_rounds) {
        if (g.counters[BSP_C_BRACKETED]) return;
        for (int p = 0; p < g.npencil; ++p)
            for (int e = 0; e < g.n; ++e) bsp_multisection_round<B>(g, p, e, r);
        if (getenv("BSP_EMUL_TRACE")) {   /* per-round census of the brackets (diagnostics) */
            int secant = 0;
            for (int p = 0; p < g.npencil; ++p)
                for (int e = 0; e < g.n; ++e) secant += g.side[(size_t)p * g.ldw + e] && !g.done[(size_t)p * g.ldw + e];
            fprintf(stderr, "round %2d open %6d crowded %6d interpolating %6d\n", r, g.counters[BSP_C_OPEN], g.counters[BSP_C_CROWDED], secant);
        }
        bsp_round_ctl(g, r, max_rounds, open_ok);
    }
    const BspSelect *sel = nullptr;
    int *nvec_eff = nullptr;
    void select() { if (sel) bsp_select_states(g, sel, nvec_eff, nullptr); }
    void prepare() {
        for (int p = 0; p < g.npencil; ++p)
            for (int e = 0; e < g.n; ++e) bsp_refine_prepare(g, p, e);
    }
    /* optional passes run on the compaction list of the previous check, like the kernels: thread `slot` works on
     * eigen index rlist[slot] and keeps its factor in column `slot` of L */
    template <class F> void for_each_active(int it, int optional, F f) {
        for (int p = 0; p < g.npencil; ++p) {
            if (optional) {
                for (int slot = 0; slot < g.n; ++slot) {
                    const int e = bsp_listed_index(g, p, slot, it);
                    if (e < 0) break;
                    f(p, e, slot);
                }
            } else {
                for (int e = 0; e < g.n; ++e) if (bsp_refine_active(g, p, e)) f(p, e, e);
            }
        }
    }
    void factor(int it, int optional) {
        cur_iter = it;
        if (optional && g.counters[BSP_C_REFINED]) return;
        for_each_active(it, optional, [&](int p, int e, int ls) {
            BspRowsGlobal<B> src{g.fbH + (size_t)p * g.nrows * (2 * B + 2), g.fbS + (size_t)g.inst[p] * g.nrows * (2 * B + 2)};
            if (ckpt && it < 2) bsp_factor_forward_rows<B, true>(g, p, e, ls, it, true, src);
            else bsp_factor_forward_rows<B>(g, p, e, ls, it, true, src);
        });
    }
    void back(int it, int cn, int cx, int optional) {
        if (optional && g.counters[BSP_C_REFINED]) return;
        for_each_active(it, optional, [&](int p, int e, int ls) {
            if (ckpt && it < 2) {
                BspRowsGlobal<B, BSP_CK_GROUPS(B)> src{g.fbH + (size_t)p * g.nrows * (2 * B + 2), g.fbS + (size_t)g.inst[p] * g.nrows * (2 * B + 2)};
                BspScratchLocal<B> scr;
                bsp_back_ckpt_rows<B>(g, p, e, ls, it, cx, true, src, scr);
            } else {
                BspRowsGlobal<B> src{g.fbH + (size_t)p * g.nrows * (2 * B + 2), g.fbS + (size_t)g.inst[p] * g.nrows * (2 * B + 2)};
                bsp_back_substitute_rows<B>(g, p, e, ls, cn, cx, true, src);
            }
        });
    }
    void resid(int optional) {
        if (optional && g.counters[BSP_C_REFINED]) return;
        for (int p = 0; p < g.npencil; ++p)
            for (int e = 0; e < g.n; ++e) bsp_residual_pass<B>(g, p, e);
    }
    void check(int it, int select) {
        if (g.counters[BSP_C_REFINED]) return;
        for (int p = 0; p < g.npencil; ++p)
            for (int e = 0; e < g.n; ++e) bsp_check_converged(g, p, e, it, select);
        if (getenv("BSP_EMUL_TRACE")) fprintf(stderr, "iteration %d unconverged %d\n", it, g.counters[BSP_C_UNCONV]);
        bsp_check_ctl(g, it);
    }
};

template <int B>
static int run(int n, int npencil, const double *hb, const double *sb, const int *nvec_in, double tau,
               double delta_rel, double conv_tol, double vec_tol, int max_rounds, int min_iters, int max_iters,
               double *E, double *C, double *stats)
{
    constexpr int K1 = B + 1, FS = 2 * B + 2;
    BspEigChunk g;
    memset(&g, 0, sizeof(g));
    g.n = n;
    g.npad = BSP_NPAD(n, B);
    g.nrows = BSP_NROWS(g.npad, B);
    g.xrows = g.npad + B + 1;
    g.ldw = ((n + 31) / 32) * 32;
    g.npencil = npencil;
    const size_t per = (size_t)npencil * g.ldw;
    std::vector<double> fbH((size_t)npencil * g.nrows * FS, 0.0), fbS((size_t)npencil * g.nrows * FS, 0.0);
    for (int p = 0; p < npencil; ++p) {
        const double *h = hb + (size_t)p * K1 * n, *s = sb + (size_t)p * K1 * n;
        double *H = fbH.data() + (size_t)p * g.nrows * FS, *S = fbS.data() + (size_t)p * g.nrows * FS;
        for (int i = 0; i < n; ++i)
            for (int d = 0; d <= B; ++d) {
                if (i + d < n) { H[(size_t)i * FS + B + d] = h[(size_t)d * n + i]; S[(size_t)i * FS + B + d] = s[(size_t)d * n + i]; }
                if (i - d >= 0) { H[(size_t)i * FS + B - d] = h[(size_t)d * n + i - d]; S[(size_t)i * FS + B - d] = s[(size_t)d * n + i - d]; }
            }
        for (int i = n; i < g.nrows; ++i) H[(size_t)i * FS + B] = 1.0;
    }
    std::vector<int> inst(npencil), nvec(npencil);
    for (int p = 0; p < npencil; ++p) { inst[p] = p; nvec[p] = nvec_in[p]; }
    std::vector<double> pbound(npencil * 4), lo(2 * per), hi(2 * per), samp_s(2 * per), gap(per), sigma(per),
        rho(per), rho_prev(per), scale(per), res(per), res2(per), xmax(per);
    std::vector<int> rlist(2 * per), rcount(2 * npencil, 0);
    std::vector<int> clo(2 * per), chi(2 * per), samp_c(2 * per), done(per), status(per), counters(BSP_C_WORDS, 0);
    std::vector<double> samp_fm(2 * per), flm(per), fhm(per), beta(per);
    std::vector<int> samp_fe(2 * per), fle(per), fhe(per), side(per);
    std::vector<double> L((size_t)npencil * g.npad * K1 * g.ldw), X((size_t)npencil * g.xrows * g.ldw, 0.0),
        R((size_t)npencil * g.xrows * g.ldw, 0.0);
    std::vector<double> CK((size_t)npencil * (g.npad / BSP_CK_STEPS(B)) * BSP_CK_DOUBLES(B) * g.ldw);
    g.fbH = fbH.data(); g.fbS = fbS.data(); g.inst = inst.data(); g.nvec = nvec.data(); g.nvec_br = nvec.data();
    g.pbound = pbound.data(); g.lo = lo.data(); g.hi = hi.data(); g.clo = clo.data(); g.chi = chi.data();
    g.samp_s = samp_s.data(); g.samp_c = samp_c.data(); g.gap = gap.data(); g.done = done.data();
    g.samp_fm = samp_fm.data(); g.samp_fe = samp_fe.data(); g.flm = flm.data(); g.fhm = fhm.data();
    g.fle = fle.data(); g.fhe = fhe.data(); g.side = side.data(); g.beta = beta.data();
    g.sigma = sigma.data(); g.rho = rho.data(); g.rho_prev = rho_prev.data(); g.scale = scale.data();
    g.res = res.data(); g.res2 = res2.data(); g.rlist = rlist.data(); g.rcount = rcount.data(); g.vec_tol = vec_tol; g.xmax = xmax.data(); g.status = status.data(); g.L = L.data(); g.X = X.data(); g.R = R.data();
    g.counters = counters.data(); g.tau = tau; g.delta_rel = delta_rel; g.conv_tol = conv_tol;
    g.CK = CK.data();
    EmulExec<B> ex; ex.g = g; ex.ckpt = getenv("BSP_EMUL_CKPT") ? atoi(getenv("BSP_EMUL_CKPT")) : 0;
    BspSchedule sch = {max_rounds, min_iters, max_iters};
    bsp_enqueue_chunk(ex, sch);
    BspRunStats st = {counters[BSP_C_ROUNDS], counters[BSP_C_ITERS], counters[BSP_C_OPEN_END], counters[BSP_C_CROWDED_END],
                      counters[BSP_C_UNCONV_END]};
    std::vector<double> fac(per);
    std::vector<int> bad(npencil, 0);
    for (int p = 0; p < npencil; ++p)
        for (int e = 0; e < n; ++e) bsp_finalize_eigen(g, p, e, E, fac.data(), bad.data(), 1e-9);
    for (int p = 0; p < npencil; ++p)
        for (int e = 0; e < nvec[p]; ++e)
            for (int j = 0; j < n; ++j)
                C[((size_t)p * n + e) * n + j] = fac[(size_t)p * g.ldw + e] * X[((size_t)p * g.xrows + j) * g.ldw + e];
    stats[0] = st.rounds; stats[1] = st.iters; stats[2] = st.brackets_open; stats[3] = st.unconverged;
    double rmax = 0;
    for (int p = 0; p < npencil; ++p) for (int e = 0; e < nvec[p]; ++e) { double r = res[(size_t)p * g.ldw + e] / fmax(1.0, fabs(rho[(size_t)p * g.ldw + e])); if (!(r <= rmax)) rmax = r; }
    stats[4] = rmax;
    int nb = 0; for (int p = 0; p < npencil; ++p) nb += bad[p];
    stats[5] = nb;
    stats[6] = counters[BSP_C_SELECTED];   /* eigenpairs that went through a third solve */
    return 0;
}

extern "C" int emul_solve(int n, int B, int npencil, const double *hb, const double *sb, const int *nvec,
                          double tau, double delta_rel, double conv_tol, double vec_tol, int max_rounds, int min_iters,
                          int max_iters, double *E, double *C, double *stats)
{
#define CASE(b) case b: return run<b>(n, npencil, hb, sb, nvec, tau, delta_rel, conv_tol, vec_tol, max_rounds, min_iters, max_iters, E, C, stats);
    switch (B) { CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) default: return -2; }
}

template <int B> static int cnt(int n, const double *hb, const double *sb, double sigma)
{
    constexpr int K1 = B + 1, FS = 2 * B + 2;
    int npad = BSP_NPAD(n, B), nrows = BSP_NROWS(npad, B);
    std::vector<double> H((size_t)nrows * FS, 0.0), S((size_t)nrows * FS, 0.0);
    for (int i = 0; i < n; ++i)
        for (int d = 0; d <= B; ++d) {
            if (i + d < n) { H[(size_t)i * FS + B + d] = hb[(size_t)d * n + i]; S[(size_t)i * FS + B + d] = sb[(size_t)d * n + i]; }
            if (i - d >= 0) { H[(size_t)i * FS + B - d] = hb[(size_t)d * n + i - d]; S[(size_t)i * FS + B - d] = sb[(size_t)d * n + i - d]; }
        }
    for (int i = n; i < nrows; ++i) H[(size_t)i * FS + B] = 1.0;
    return bsp_sturm_count<B>(H.data(), S.data(), npad, sigma, 1e-300, nullptr);
}
extern "C" int emul_count(int n, int B, const double *hb, const double *sb, double sigma)
{
#define CASE2(b) case b: return cnt<b>(n, hb, sb, sigma);
    switch (B) { CASE2(1) CASE2(2) CASE2(3) CASE2(4) CASE2(5) CASE2(6) CASE2(7) CASE2(8) CASE2(9) default: return -2; }
}
