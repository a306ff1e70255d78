/*
 * bsp_core.h -- per-thread bodies of the banded generalized eigensolver.
 *
 * One CUDA thread owns one eigenpair (pencil p, index e) and walks the band of
 * H - sigma S row by row with a (B+1)x(B+1) Schur-complement window kept in
 * registers (all indices static after unrolling by B+1).  The bodies are
 * written as host/device functions over plain pointers so the __global__
 * wrappers in bsp_kernels.cu stay thin, and so tests/ can replay exactly this
 * logic on the CPU (tests/emul/; test infrastructure, never a product path).
 *
 * Replaces the work LAPACK DSYGV does at matrices.f90:248 of the reference
 * (dpotrf -> dsygst -> dsytrd -> dorgtr -> dsteqr -> dtrsm, all dense O(N^3))
 * by a band-native pipeline, O(N^2 b^2) per pencil and parallel over the N
 * eigenpairs:
 *   1. multisection with Sturm counts  nu(sigma) = #negative pivots of the
 *      banded LDL^T of H - sigma S  (Sylvester inertia; S is SPD),
 *   2. inverse iteration / Rayleigh-quotient iteration on the banded pencil,
 *      the last steps in residual-correction form so that the un-pivoted LDL^T
 *      only has to be a contraction, not an accurate solver,
 *   3. S-normalisation, sign convention, transpose to column-major C.
 *
 * Band storage ("full-band rows"): row i of a matrix with half bandwidth B is
 *   fb[i*FS + c] = A(i, i-B+c),  c = 0..2B,  FS = 2B+2  (0-based i).
 * Rows n..nrows-1 are padding: diag(H)=1, everything else 0, which decouples
 * them and keeps every count unchanged.
 */
#ifndef BSP_CORE_H
#define BSP_CORE_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define BSP_HD __host__ __device__ __forceinline__
#else
#define BSP_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define BSP_LDG(p) __ldg(p)
/* reciprocal of a pivot: the arguments are finite, normal and bounded away from zero (|d| >= pivmin), so the
 * special-case path of __drcp_rn (a branch + call per row of every sweep) is not needed: hardware seed
 * (MUFU.RCP64H, ~20 bits), one cubic and one linear correction, the same arithmetic as the library's fast path */
__device__ __forceinline__ double bsp_drcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
#define BSP_RCP(x) bsp_drcp(x)
#else
#define BSP_LDG(p) (*(p))
#define BSP_RCP(x) (1.0 / (x))
#endif

#define BSP_EPS 2.220446049250313e-16

/* -DBSP_DEBUG (libbspatom_debug.so, tests/test_gpu_debug.py): bounds checks on every workspace / list index and on the
 * tile pipeline's bookkeeping as device-side asserts -- compute-sanitizer is closed on this pool, so this build is the
 * memory-safety check of the kernels.  A failed assert traps: the host sees a CUDA error, never a silent wrong answer. */
#if defined(BSP_DEBUG)
#include <assert.h>
#define BSP_ASSERT(c) assert(c)
#else
#define BSP_ASSERT(c) ((void)0)
#endif
#define BSP_ABS_CLOSE 2e-14   /* absolute bracket width (Hartree) at which a bracket counts as closed */

/* rows stored per band matrix: the sweeps bring in row j+B+1 at step j and
 * prefetch one row further, so npad + B + 2 rows exist (padding rows are
 * decoupled: diag(H) = 1, rest 0) */
#define BSP_NROWS(npad, B) ((npad) + (B) + 2)

/* rows are padded to a whole number of BSP_SEG_BLOCKS groups of B+1 steps (a whole number of tiles) */
#ifndef BSP_SEG_BLOCKS
#define BSP_SEG_BLOCKS 4
#endif
#define BSP_SEG_STEPS(B) (BSP_SEG_BLOCKS * ((B) + 1))
#define BSP_NPAD(n, B) ((((n) + BSP_SEG_STEPS(B) - 1) / BSP_SEG_STEPS(B)) * BSP_SEG_STEPS(B))

/* Check-pointed solves (the two full-width solves of every eigenpair): the forward sweep does not store the
 * factor (B+1 doubles per row and eigenpair: 57 MB per pencil at N = 1000, written once and read once) but, at
 * the start of every segment of BSP_CK_GROUPS(B) groups of B+1 rows, the part of the elimination state that
 * cannot be re-read from the band: rows 0..B-1 of the pivot window (lower triangle), the right-hand-side window
 * y[0..B] and the B+1 raw right-hand-side values already in flight -- BSP_CK_DOUBLES(B) doubles.  The back sweep
 * takes the segments last to first: it re-eliminates a segment from its check-point into a scratch of
 * BSP_CK_STEPS(B) x (B+1) doubles (shared memory in the kernel: the thread-private local-memory scratch of round 1
 * thrashed L1) and back-substitutes it at once.  HBM traffic of the pair of sweeps at
 * B = 6: 2 x 56 -> 2 x 20 bytes per row and eigenpair, at the price of eliminating twice (the FP64 pipe idles at
 * 17-26 % in the stored-factor sweeps). */
#define BSP_CK_GROUPS(B) ((B) <= 6 ? 2 : 1)
#define BSP_CK_STEPS(B) (BSP_CK_GROUPS(B) * ((B) + 1))
#define BSP_CK_DOUBLES(B) ((((B)) * ((B) + 1)) / 2 + 2 * ((B) + 1))

/* The sweeps walk the band rows of a pencil in tiles of BSP_TILE_STEPS(B) steps (whole unrolled groups of
 * B+1 steps; npad is a whole number of tiles).  A row source hands out one tile at a time: on the host (and
 * for the bounds ladder) straight from global memory, in the round / factor / back kernels from the
 * shared-memory stages the block fills with bulk copies (BspRowsStaged in bsp_kernels.cuh). */
#define BSP_TILE_GROUPS(B) ((B) <= 8 ? 4 : 2)
#define BSP_TILE_STEPS(B) (BSP_TILE_GROUPS(B) * ((B) + 1))

/* refinement status bits */
#define BSP_ST_CONVERGED 1
#define BSP_F_UNKNOWN (-2000000000)

/* device-side control block of a chunk (BspEigChunk::counters): the whole stage schedule of a chunk is
 * enqueued without any host read; kernels that are no longer needed see a flag and return at once */
#define BSP_C_OPEN 0        /* brackets still open after the current round                      */
#define BSP_C_UNCONV 1      /* eigenpairs above conv_tol after the current iteration            */
#define BSP_C_CROWDED 2     /* open brackets that do not isolate one eigenvalue yet             */
#define BSP_C_BRACKETED 3   /* flag: bracketing finished, later round kernels are no-ops        */
#define BSP_C_BUF 4         /* bracket buffer (0/1) holding the final brackets                  */
#define BSP_C_ROUNDS 5      /* rounds executed                                                  */
#define BSP_C_REFINED 6     /* flag: every eigenpair converged, later iterations are no-ops     */
#define BSP_C_ITERS 7       /* refinement iterations executed                                   */
#define BSP_C_ARRIVE 9      /* block arrival counter of the "last block does the bookkeeping"   */
#define BSP_C_OPEN_END 10   /* open brackets at hand-over                                       */
#define BSP_C_CROWDED_END 11
#define BSP_C_UNCONV_END 12
#define BSP_C_SELECTED 13    /* eigenpairs the first convergence check kept in the iteration     */
#define BSP_C_WORDS 16

struct BspEigChunk {
    /* geometry */
    int n;        /* basis size                                   */
    int npad;     /* BSP_NPAD(n, B): whole segments of (B+1) blocks */
    int nrows;    /* rows stored per band matrix = BSP_NROWS      */
    int xrows;    /* rows of X / R workspaces   = npad + B + 1    */
    int ldw;      /* eigen-index stride (n rounded up to 32)      */
    int npencil;  /* pencils in this chunk                        */
    /* pencils */
    const double *fbH; /* [npencil][nrows][FS]  H_l = H0 + c_l Q  */
    const double *fbS; /* [ninst][nrows][FS]                      */
    const int *inst;   /* [npencil] instance of each pencil       */
    const int *nvec;   /* [npencil] eigenvectors wanted (with device-side selection: written by bsp_select_states
                          after the bracketing, <= the caller's cap)                                         */
    const int *nvec_br; /* [npencil] the same as seen by the bracketing: the caller's nvec, or zeros when the
                          selection needs every eigenvalue to rounding before it can decide */
    double *pbound;    /* [npencil][4]  lo0, hi0, hmax, smax      */
    /* bracket state, double buffered: index (buf*npencil + p)*ldw + e */
    double *lo, *hi;
    int *clo, *chi;
    double *samp_s; /* [2][npencil][ldw] published samples        */
    int *samp_c;
    double *samp_fm; /* [2][npencil][ldw] det(H - s S) of the sample: */
    int *samp_fe;    /*   mantissa in +-[0.5,1) and binary exponent   */
    /* private per eigen index [npencil][ldw]: det at the bracket ends
     * (exponent BSP_F_UNKNOWN = not evaluated) and the Illinois side flag */
    double *flm, *fhm;
    int *fle, *fhe, *side;
    double *beta; /* bracket width at the previous sample (stall detection) */
    double *gap;    /* [npencil][ldw] lower bound of the gap      */
    int *done;      /* [npencil][ldw]                             */
    /* refinement state [npencil][ldw] */
    double *sigma, *rho, *rho_prev, *scale, *res;
    double *res2;   /* 2-norm of the residual of the S-normalised vector (selects who gets a correction pass) */
    double *xmax;   /* max |x_j| of the vector in X (tracked by the back sweep; the sign convention needs it) */
    int *status;
    /* workspaces */
    double *L; /* [npencil][npad][B+1][ldw]  (zd, l_1..l_B): stored factor (compacted correction passes) */
    double *CK; /* [npencil][npad / BSP_CK_STEPS][BSP_CK_DOUBLES][ldw]: check-points of the full-width solves,
                   or null: every solve stores its factor */
    double *X; /* [npencil][xrows][ldw]                           */
    double *R; /* [npencil][xrows][ldw]                           */
    int *counters; /* control block, BSP_C_* */
    /* compaction of the bracketing rounds (device only, may be null): olist[buf][p][.] = eigen indices of
     * pencil p that still have work in the round reading buffer buf -- open brackets from the front,
     * just-finished ones from the back; ocount[buf][p][2] = how many of each */
    int *olist, *ocount;
    /* compaction of the optional refinement passes (may be null: every pass then runs at full width):
     * rlist[buf][p][.] = eigen indices of pencil p that are not converged yet, rcount[buf][p] how many.
     * Written by the convergence check of iteration t into buffer t & 1 and read by the sweeps of iteration
     * t + 1, whose thread `slot` works on eigen index rlist[slot] and keeps its factor in column `slot` of L:
     * the factor stream (7/9 of the sweep traffic) stays fully coalesced however sparse the active set is. */
    int *rlist, *rcount;
    /* tunables */
    double tau;       /* bracket width / gap at hand-over         */
    double delta_rel; /* shift offset / gap in correction steps   */
    double conv_tol;  /* scaled residual that ends the iteration  */
    double vec_tol;   /* ||r||_2 / gap above which the second solve is followed by a correction pass */
};

/* ------------------------------------------------------------------------- */
BSP_HD double bsp_hash_uniform(uint32_t a, uint32_t b, uint32_t c)
{
    /* stateless integer hash -> uniform in (-1, 1); start vector of the
     * inverse iteration (dstein draws a random vector at the same point) */
    uint32_t x = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return ((double)x + 0.5) * (2.0 / 4294967296.0) - 1.0;
}

/* ------------------------------------------------------------------------- *
 * Sturm count: number of negative pivots of LDL^T(H - sigma S) over rows
 * 0..npad-1.  first_neg (optional) receives the first row with a pivot <= 0.
 * ------------------------------------------------------------------------- */
BSP_HD void bsp_renorm(double &m, int &e)
{
    /* m <- m / 2^k with |m| in [0.5, 1), e += k  (m finite, non-zero, normal) */
#if defined(__CUDA_ARCH__)
    long long b = __double_as_longlong(m);
#else
    long long b;
    memcpy(&b, &m, sizeof b);
#endif
    const int ex = (int)((b >> 52) & 0x7ff) - 1022;
    b = (b & (long long)0x800fffffffffffffULL) | (long long)0x3fe0000000000000ULL;
#if defined(__CUDA_ARCH__)
    m = __longlong_as_double(b);
#else
    memcpy(&m, &b, sizeof b);
#endif
    e += ex;
}

/* The sweep is written as "window initialisation" + "group of B+1 steps" so that the same arithmetic runs
 * from global memory (GL = true: host replay, bounds ladder) and from the shared-memory tiles the round
 * kernel stages with bulk copies (GL = false).  rows point at band rows of FS doubles. */
template <bool GL>
BSP_HD double bsp_ld(const double *p)
{
#if defined(__CUDA_ARCH__)
    if (GL) return __ldg(p);
#endif
    return *p;
}

template <int B, bool GL>
BSP_HD void bsp_sturm_init(double (&w)[B + 1][B + 1], const double *__restrict__ rowsH,
                           const double *__restrict__ rowsS, double sigma)
{
    constexpr int K1 = B + 1;
    constexpr int FS = 2 * B + 2;
#pragma unroll
    for (int r = 0; r < K1; ++r) {
#pragma unroll
        for (int c = 0; c < K1; ++c) {
            if (c <= r) {
                const int off = r * FS + (c - r + B);
                w[r][c] = fma(-sigma, bsp_ld<GL>(rowsS + off), bsp_ld<GL>(rowsH + off));
            } else {
                w[r][c] = 0.0;
            }
        }
    }
}

/* steps j0 .. j0+B; hnext / snext = band row j0 + B + 1, the row that enters the window at the end of step
 * j0.  It is read at the top of its step: from a staged tile the shared-memory latency hides behind the pivot
 * arithmetic of the step, and no row is carried in registers across steps. */
template <int B, bool GL>
BSP_HD void bsp_sturm_group(double (&w)[B + 1][B + 1], const double *__restrict__ hnext,
                            const double *__restrict__ snext, double sigma, double pivmin, int j0, int &cnt,
                            int &first, double &fm, int &fe)
{
    constexpr int K1 = B + 1;
    constexpr int FS = 2 * B + 2;
#pragma unroll
    for (int t = 0; t < K1; ++t) {
        double nh[K1], ns[K1];
#pragma unroll
        for (int m = 0; m <= B; ++m) {
            nh[m] = bsp_ld<GL>(hnext + (size_t)t * FS + m);
            ns[m] = bsp_ld<GL>(snext + (size_t)t * FS + m);
        }
        double d = w[t][t];
        if (fabs(d) < pivmin) d = -pivmin;
        if (d < 0.0) { ++cnt; if (first < 0) first = j0 + t; }
        /* det = product of the pivots, kept as mantissa * 2^fe; |d| >= pivmin and a huge pivot only follows a
         * tiny one, so the mantissa may run over four rows before it is brought back to [0.5, 1) */
        fm *= d;
        if ((t & 3) == 3 || t == K1 - 1) bsp_renorm(fm, fe);
        const double rinv = BSP_RCP(d);
        double col[K1], l[K1];
#pragma unroll
        for (int i = 1; i <= B; ++i) {
            col[i] = w[(t + i) % K1][t];
            l[i] = col[i] * rinv;
        }
#pragma unroll
        for (int m = 1; m <= B; ++m) {
#pragma unroll
            for (int i = m; i <= B; ++i) {
                w[(t + i) % K1][(t + m) % K1] = fma(-l[i], col[m], w[(t + i) % K1][(t + m) % K1]);
            }
        }
        /* row j retires; its slot takes row j+B+1 */
#pragma unroll
        for (int m = 0; m <= B; ++m) w[t][(t + 1 + m) % K1] = fma(-sigma, ns[m], nh[m]);
    }
}

struct BspTrue { static constexpr bool value = true; };
struct BspFalse { static constexpr bool value = false; };

/* row source reading global memory.  Forward tiles: pointer to band row t*TR, rows t*TR .. (t+1)*TR+B are
 * read.  Backward tiles (taken in descending t): same pointer, rows t*TR .. (t+1)*TR-1 are read. */
template <int B, int G = BSP_TILE_GROUPS(B)>
struct BspRowsGlobal {
    static constexpr bool GL = true;
    static constexpr int RHS_RING = 0;   /* no shared-memory ring for the right-hand side */
    static constexpr int TR = G * (B + 1);   /* steps per tile */
    const double *H, *S;
    BSP_HD void begin_forward(int) {}
    BSP_HD void acquire_forward(int t, const double *&tH, const double *&tS)
    {
        tH = H + (size_t)t * TR * (2 * B + 2);
        tS = S + (size_t)t * TR * (2 * B + 2);
    }
    BSP_HD void release_forward(int, int) {}
    BSP_HD void begin_backward(int) {}
    BSP_HD void acquire_backward(int t, int, const double *&tH, const double *&tS) { acquire_forward(t, tH, tS); }
    BSP_HD void release_backward(int, int) {}
    /* backward order, forward-sized tiles (rows t*TR .. (t+1)*TR + B): the check-pointed back sweep */
    BSP_HD void begin_backward_wide(int) {}
    BSP_HD void acquire_backward_wide(int t, int, const double *&tH, const double *&tS) { acquire_forward(t, tH, tS); }
    BSP_HD void release_backward_wide(int, int) {}
};

/* called by every thread that shares `src` (the staged source synchronises the block);
 * only `active` threads do arithmetic */
template <int B, class Src>
BSP_HD int bsp_sturm_sweep(Src &src, int npad, bool active, double sigma, double pivmin, int *first_neg,
                           double *det_m, int *det_e)
{
    constexpr int K1 = B + 1;
    constexpr int FS = 2 * B + 2;
    constexpr int TR = Src::TR;
    const int ntiles = npad / TR;
    double w[K1][K1];
    int cnt = 0, first = -1;
    double fm = 0.5; /* det(H - sigma S) = prod of pivots = fm * 2^fe */
    int fe = 1;
    src.begin_forward(ntiles);
    for (int t = 0; t < ntiles; ++t) {
        const double *tH, *tS;
        src.acquire_forward(t, tH, tS);
        if (active) {
            if (t == 0) bsp_sturm_init<B, Src::GL>(w, tH, tS, sigma);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int gq = 0; gq < TR; gq += K1)
                bsp_sturm_group<B, Src::GL>(w, tH + (size_t)(gq + K1) * FS, tS + (size_t)(gq + K1) * FS, sigma, pivmin,
                                            t * TR + gq, cnt, first, fm, fe);
        }
        src.release_forward(t, ntiles);
    }
    if (first_neg) *first_neg = first;
    if (det_m) { *det_m = fm; *det_e = fe; }
    return cnt;
}

template <int B>
BSP_HD int bsp_sturm_count(const double *__restrict__ fbH, const double *__restrict__ fbS, int npad,
                           double sigma, double pivmin, int *first_neg, double *det_m = nullptr,
                           int *det_e = nullptr)
{
    BspRowsGlobal<B> src{fbH, fbS};
    return bsp_sturm_sweep<B>(src, npad, true, sigma, pivmin, first_neg, det_m, det_e);
}

/* ------------------------------------------------------------------------- *
 * bounds: 64 candidates per pencil around s0 = max_j |H_jj| / S_jj:
 *   t = 0..31  : +s0 * 2^(t-8)      (2^-8 .. 2^23)
 *   t = 32..63 : -s0 * 2^(t-32-16)  (2^-16 .. 2^15)
 * pbound[4p..4p+3] = lo0, hi0, hmax, smax ; lo0 > hi0 flags "not bracketed".
 * ------------------------------------------------------------------------- */
#define BSP_NCAND 64
/* shift of candidate `lane` of pencil p (and the pencil's scale figures, stored by lane 0) */
template <int B>
BSP_HD void bsp_bounds_shift(const BspEigChunk &g, int p, int lane, double &sig, double &pivmin)
{
    constexpr int FS = 2 * B + 2;
    const double *fbH = g.fbH + (size_t)p * g.nrows * FS;
    const double *fbS = g.fbS + (size_t)g.inst[p] * g.nrows * FS;
    double s0 = 0.0, hmax = 0.0, smax = 0.0;
    for (int j = 0; j < g.n; ++j) {
        const double h = fabs(fbH[(size_t)j * FS + B]);
        const double s = fbS[(size_t)j * FS + B];
        if (h > hmax) hmax = h;
        if (s > smax) smax = s;
        const double q = h / s;
        if (q > s0) s0 = q;
    }
    if (!(s0 > 0.0) || !(s0 < INFINITY)) s0 = 1.0;
    if (lane < 32) sig = ldexp(s0, lane - 8);
    else sig = -ldexp(s0, lane - 32 - 16);
    pivmin = 1e-30 * (hmax + fabs(sig) * smax);
    if (lane == 0) {
        g.pbound[p * 4 + 2] = hmax;
        g.pbound[p * 4 + 3] = smax;
    }
}

template <int B>
BSP_HD void bsp_bounds_candidate(const BspEigChunk &g, int p, int lane, double *cand_s, int *cand_c)
{
    constexpr int FS = 2 * B + 2;
    double sig, pivmin;
    bsp_bounds_shift<B>(g, p, lane, sig, pivmin);
    cand_s[p * BSP_NCAND + lane] = sig;
    cand_c[p * BSP_NCAND + lane] = bsp_sturm_count<B>(g.fbH + (size_t)p * g.nrows * FS, g.fbS + (size_t)g.inst[p] * g.nrows * FS,
                                                      g.npad, sig, pivmin, nullptr);
}

BSP_HD void bsp_bounds_pick(const BspEigChunk &g, int p, const double *cand_s, const int *cand_c)
{
    /* hi0: smallest positive candidate with nu = n; lo0: negative candidate of
     * smallest magnitude with nu = 0.  If a ladder never brackets the spectrum
     * lo0 > hi0 is stored and every eigenpair of the pencil is reported bad. */
    double hi0 = -1.0, lo0 = 1.0;
    for (int t = 31; t >= 0; --t)
        if (cand_c[p * BSP_NCAND + t] >= g.n) hi0 = cand_s[p * BSP_NCAND + t];
    for (int t = 63; t >= 32; --t)
        if (cand_c[p * BSP_NCAND + t] == 0) lo0 = cand_s[p * BSP_NCAND + t];
    g.pbound[p * 4 + 0] = lo0;
    g.pbound[p * 4 + 1] = hi0;
}

/* ------------------------------------------------------------------------- *
 * one bracketing round for eigen index e of pencil p.
 * reads bracket buffer (round&1), writes buffer ((round+1)&1).
 *
 * While a bracket holds m > 1 eigenvalues its owners split it evenly
 * (multisection).  Once it isolates eigenvalue e (m == 1) and det(H - sigma S)
 * is known at both ends (the product of the pivots comes for free with the
 * count), the sample is the regula-falsi point of the determinant, pushed to
 * twice its distance from the nearer end when it hugs that end: pairs of
 * rounds then square the bracket width (order ~1.41 per round) instead of one
 * bit per round.  The inertia of every sample keeps the bracket rigorous.
 * ------------------------------------------------------------------------- */
/* state of one eigen index between the phases of a round */
struct BspRoundState {
    double lo, hi, flm, fhm, beta, gp, wdt, frac, s, sfm;
    int clo, chi, fle, fhe, side, done, c, sfe;
    int was_done;   /* done before this round started (the round then only re-publishes its state) */
    int nfail;      /* 1: the previous sample was a middle estimate that did not halve the bracket */
    int want_defl;  /* the regula-falsi step is possible: the deflation sum at the bracket midpoint is needed */
    int want_count; /* s is a new sample: its inertia is needed                                              */
};

/* phase 1: load the bracket, tighten it with last round's samples, decide done / kind of step */
BSP_HD void bsp_round_begin(const BspEigChunk &g, int p, int e, int round, BspRoundState &st)
{
    const int n = g.n;
    const size_t per = (size_t)g.npencil * g.ldw;
    const size_t id = (size_t)p * g.ldw + e;
    const size_t rd = (size_t)(round & 1) * per;
    double lo, hi, flm = 0.0, fhm = 0.0, beta = 0.0;
    int clo, chi, fle = BSP_F_UNKNOWN, fhe = BSP_F_UNKNOWN, side = 0; /* side >= 1: last sample was an interpolated one */
    if (round == 0) {
        lo = g.pbound[p * 4 + 0]; hi = g.pbound[p * 4 + 1]; clo = 0; chi = n;
        if (lo > hi) { const double t_ = lo; lo = hi; hi = t_; } /* unbracketed: flagged in finalize */
    } else {
        lo = g.lo[rd + id]; hi = g.hi[rd + id]; clo = g.clo[rd + id]; chi = g.chi[rd + id];
        flm = g.flm[id]; fle = g.fle[id]; fhm = g.fhm[id]; fhe = g.fhe[id]; side = g.side[id]; beta = g.beta[id];
    }
    const int was_done = (round == 0) ? 0 : g.done[id];
    if (round > 0 && !was_done) {
        /* tighten with the samples every eigen index of this pencil published
         * last round: s is non-decreasing in the index, c = nu(s) monotone */
        const size_t so = (size_t)((round - 1) & 1) * per + (size_t)p * g.ldw;
        const double *S = g.samp_s + so;
        const int *Cc = g.samp_c + so;
        int a = -1, b = n;
        while (b - a > 1) {
            const int mid = (a + b) >> 1;
            if (Cc[mid] > e) b = mid; else a = mid;
        }
        if (a >= 0) {
            const double s = S[a];
            if (s > lo && s < hi) { lo = s; clo = Cc[a]; flm = g.samp_fm[so + a]; fle = g.samp_fe[so + a]; }
        }
        if (b < n) {
            const double s = S[b];
            if (s < hi && s > lo) { hi = s; chi = Cc[b]; fhm = g.samp_fm[so + b]; fhe = g.samp_fe[so + b]; }
        }
    }
    /* gap to the neighbours' brackets (theirs as of last round: still valid) */
    double gl = INFINITY, gr = INFINITY;
    if (round > 0) {
        if (e > 0) gl = lo - g.hi[rd + id - 1];
        if (e + 1 < n) gr = g.lo[rd + id + 1] - hi;
    } else {
        if (n > 1) { gl = 0.0; gr = 0.0; }
    }
    const double gp = fmin(gl, gr);
    const double wdt = hi - lo;
    const double amax = fmax(fabs(lo), fabs(hi));
    int done = was_done;
    if (!done) {
        /* closed to rounding; near-zero levels to BSP_ABS_CLOSE Hartree (north star: 1e-10 absolute for near-zero levels):
         * relative to |E| ~ 1e-4 the counts are noise long before 4 eps |E|, and the bracket would creep for 60+ rounds */
        if (wdt <= fmax(4.0 * BSP_EPS * amax, BSP_ABS_CLOSE)) done = 1;
        else if (e < g.nvec_br[p] && gp > 0.0 && wdt <= g.tau * gp) done = 1;
    }
    st.lo = lo; st.hi = hi; st.flm = flm; st.fhm = fhm; st.beta = beta; st.gp = gp; st.wdt = wdt;
    st.clo = clo; st.chi = chi; st.fle = fle; st.fhe = fhe; st.side = side; st.done = done; st.was_done = was_done;
    st.s = lo; st.sfm = flm; st.c = clo; st.sfe = fle;
    st.want_defl = 0; st.want_count = 0; st.frac = 0.5; st.nfail = 0;
    if (!done) {
        int m = chi - clo, rk = e - clo;
        if (m < 1) m = 1;
        if (rk < 0) rk = 0;
        if (rk > m - 1) rk = m - 1;
        double frac = ((double)rk + 0.5) / (double)m;
        if (round == 0) frac = frac * frac; /* box states: E_i ~ i^2 */
        st.frac = frac;
        /* beta holds the bracket width at the previous interpolated sample; side says what kind it was:
         * 1 = an overshooting sample (the estimate hugged one end and the sample was put at twice its distance:
         * the bracket is expected to collapse), 2 = the estimate itself, somewhere in the middle (it lands on
         * one side of the root and leaves the far end where it was: the bracket halves at best, and the next,
         * overshooting, sample closes it from the other side).  An overshooting sample that did not halve the
         * bracket means the model is off: this round bisects (Brent-style safeguard against creeping).  After a
         * middle sample that did not halve it, only an overshooting sample is accepted (see bsp_round_pick). */
        const bool failed = (side >= 1) && (wdt > 0.5 * beta);
        const bool stalled = failed && side == 1;
        st.nfail = (failed && side == 2) ? 1 : 0;
        if (m == 1 && round > 0 && !stalled && fle != BSP_F_UNKNOWN && fhe != BSP_F_UNKNOWN &&
            ((flm < 0.0) != (fhm < 0.0)))
            st.want_defl = 1;
    }
}

/* phase 2 (plain form; the round kernel computes the same sum from shared-memory tiles):
 * det(H - sigma S) = prod_k (lambda_k - sigma) varies over the bracket like
 * (lambda_e - sigma) * exp(beta sigma), beta = sum_{k != e} 1/(sigma - lambda_k) (hundreds of
 * clustered levels make |beta| w >> 1).  Deflate that factor with the other brackets'
 * current midpoints (Maehly deflation; float reciprocals are plenty for a slope). */
BSP_HD double bsp_deflation_sum(const BspEigChunk &g, int p, int e, int round, const BspRoundState &st)
{
    const int n = g.n;
    const size_t rd = (size_t)(round & 1) * (size_t)g.npencil * g.ldw;
    const double mid = 0.5 * (st.lo + st.hi);
    const double *Lo = g.lo + rd + (size_t)p * g.ldw, *Hi = g.hi + rd + (size_t)p * g.ldw;
    float b0 = 0.0f, b1 = 0.0f, b2 = 0.0f, b3 = 0.0f; /* independent chains: loads overlap */
    int k = 0;
    for (; k + 4 <= n; k += 4) {
        const float d0 = (float)(mid - 0.5 * (Lo[k] + Hi[k]));
        const float d1 = (float)(mid - 0.5 * (Lo[k + 1] + Hi[k + 1]));
        const float d2 = (float)(mid - 0.5 * (Lo[k + 2] + Hi[k + 2]));
        const float d3 = (float)(mid - 0.5 * (Lo[k + 3] + Hi[k + 3]));
        b0 += (k != e && d0 != 0.0f) ? 1.0f / d0 : 0.0f;
        b1 += (k + 1 != e && d1 != 0.0f) ? 1.0f / d1 : 0.0f;
        b2 += (k + 2 != e && d2 != 0.0f) ? 1.0f / d2 : 0.0f;
        b3 += (k + 3 != e && d3 != 0.0f) ? 1.0f / d3 : 0.0f;
    }
    for (; k < n; ++k) {
        const float dk = (float)(mid - 0.5 * (Lo[k] + Hi[k]));
        b0 += (k != e && dk != 0.0f) ? 1.0f / dk : 0.0f;
    }
    return (double)b0 + (double)b1 + (double)b2 + (double)b3;
}

/* phase 3: choose the sample */
BSP_HD void bsp_round_pick(BspRoundState &st, double bsum)
{
    if (st.done) return;
    double frac = st.frac;
    bool secant = false;
    int kind = 0;
    if (st.want_defl) {
        /* regula falsi on the deflated determinant: root at lo + w / (1 + r),
         * r = |f(hi)/f(lo)| exp(-beta w) */
        const double shift = -bsum * st.wdt * 1.4426950408889634; /* in powers of two */
        double de = (double)(st.fhe - st.fle) + shift;
        de = de > 1000.0 ? 1000.0 : (de < -1000.0 ? -1000.0 : de);
        const double dei = floor(de);
        const double r = ldexp(fabs(st.fhm / st.flm) * exp2(de - dei), (int)dei);
        const double t = 1.0 / (1.0 + r);
        /* the estimate sits at fraction t; when it hugs one end, sample at twice its distance
         * from that end: the root then (almost surely) lies between the end and the sample and
         * the bracket collapses to ~2x the interpolation error instead of creeping one-sidedly */
        if (t > 0.0 && t < 1.0) {
            if (t < 0.25 || t > 0.75) {
                frac = t < 0.25 ? 2.0 * t : 1.0 - 2.0 * (1.0 - t);
                secant = true;
                kind = 1;
            } else if (!st.nfail) {
                frac = t;
                secant = true;
                kind = 2;
            }   /* else: a second middle estimate in a row after a poor one: bisect */
        }
    }
    double s = st.lo + st.wdt * frac;
    if (secant) {
        /* an estimate that is numerically ON one end (root within an ulp of it: the other end then crept towards it
         * by bisection, one bit per round, for 20+ rounds) is sampled two ulps inside instead: one count closes
         * the bracket to rounding */
        const double tiny = 2.0 * BSP_EPS * fmax(fabs(st.lo), fabs(st.hi));
        if (s - st.lo < tiny) s = st.lo + tiny;
        if (st.hi - s < tiny) s = st.hi - tiny;
    }
    if (!(s > st.lo && s < st.hi)) { s = st.lo + 0.5 * st.wdt; secant = false; }
    st.side = secant ? kind : 0;
    st.beta = st.wdt;
    if (!(s > st.lo && s < st.hi)) {
        st.done = 1; st.s = st.lo; st.c = st.clo;
    } else {
        st.s = s;
        st.want_count = 1;
    }
}

BSP_HD double bsp_round_pivmin(const BspEigChunk &g, int p, double s)
{
    return 1e-30 * (g.pbound[p * 4 + 2] + fabs(s) * g.pbound[p * 4 + 3]);
}

/* phase 5: fold the inertia of the sample (st.c, st.sfm, st.sfe) into the bracket and publish */
BSP_HD void bsp_round_end(const BspEigChunk &g, int p, int e, int round, BspRoundState &st)
{
    const size_t per = (size_t)g.npencil * g.ldw;
    const size_t id = (size_t)p * g.ldw + e;
    const size_t wr = (size_t)((round + 1) & 1) * per;
    if (st.want_count) {
        if (st.c <= e) {
            st.lo = st.s; st.clo = st.c; st.flm = st.sfm; st.fle = st.sfe;
        } else {
            st.hi = st.s; st.chi = st.c; st.fhm = st.sfm; st.fhe = st.sfe;
        }
    }
#if defined(BSP_TRACE) && !defined(__CUDA_ARCH__)
    if (round >= BSP_TRACE && !st.done) printf("r%d e%d lo=%.17g hi=%.17g w=%.3e gp=%.3e clo=%d chi=%d fl=(%g,%d) fh=(%g,%d) s=%.17g c=%d\n", round, e, st.lo, st.hi, st.hi-st.lo, st.gp, st.clo, st.chi, st.flm, st.fle, st.fhm, st.fhe, st.s, st.c);
#endif
    g.lo[wr + id] = st.lo; g.hi[wr + id] = st.hi; g.clo[wr + id] = st.clo; g.chi[wr + id] = st.chi;
    g.flm[id] = st.flm; g.fle[id] = st.fle; g.fhm[id] = st.fhm; g.fhe[id] = st.fhe; g.side[id] = st.side; g.beta[id] = st.beta;
    const size_t po = (size_t)(round & 1) * per + id;
    g.samp_s[po] = st.s;
    g.samp_c[po] = st.c;
    g.samp_fm[po] = st.sfm;
    g.samp_fe[po] = st.sfe;
    g.gap[id] = st.gp;
    g.done[id] = st.done;
    if (!st.done) {
        /* counters[0]: brackets still open; counters[2]: open AND not yet isolating one eigenvalue.  An eigen
         * index whose vector is not wanted (e >= nvec) is never touched by the refinement: its bracket must
         * close here, so it counts as crowded until it does (no hand-over while any is open). */
        const int crowded = (st.chi - st.clo != 1 || !(st.gp > 0.0) || e >= g.nvec_br[p]) ? 1 : 0;
#if defined(__CUDA_ARCH__)
        atomicAdd(g.counters + BSP_C_OPEN, 1);
        if (crowded) atomicAdd(g.counters + BSP_C_CROWDED, 1);
#else
        g.counters[BSP_C_OPEN] += 1;
        g.counters[BSP_C_CROWDED] += crowded;
#endif
    }
}

/* the whole round of one eigen index from global memory (host replay; the round kernel runs the same
 * phases with block-cooperative shared-memory staging in phases 2 and 4) */
template <int B>
BSP_HD void bsp_multisection_round(const BspEigChunk &g, int p, int e, int round)
{
    constexpr int FS = 2 * B + 2;
    if (e >= g.n) return;
    BspRoundState st;
    bsp_round_begin(g, p, e, round, st);
    double bsum = 0.0;
    if (st.want_defl) bsum = bsp_deflation_sum(g, p, e, round, st);
    bsp_round_pick(st, bsum);
    if (st.want_count) {
        const double *fbH = g.fbH + (size_t)p * g.nrows * FS;
        const double *fbS = g.fbS + (size_t)g.inst[p] * g.nrows * FS;
        st.c = bsp_sturm_count<B>(fbH, fbS, g.npad, st.s, bsp_round_pivmin(g, p, st.s), nullptr, &st.sfm, &st.sfe);
    }
    bsp_round_end(g, p, e, round, st);
}

/* bookkeeping after a bracketing round (one thread, after every eigen index of the chunk has run):
 * decides whether the bracketing is finished.  open_ok stragglers may be handed over open provided each
 * isolates its eigenvalue: the refinement keeps bracketing with the inertia of its own factorisations. */
BSP_HD void bsp_round_ctl(const BspEigChunk &g, int round, int max_rounds, int open_ok)
{
    int *c = g.counters;
    if (c[BSP_C_BRACKETED]) return;
    const int open = c[BSP_C_OPEN], crowded = c[BSP_C_CROWDED];
    c[BSP_C_ROUNDS] = round + 1;
    if (open == 0 || (open <= open_ok && crowded == 0) || round + 1 >= max_rounds) {
        c[BSP_C_BRACKETED] = 1;
        c[BSP_C_BUF] = (round + 1) & 1;
        c[BSP_C_OPEN_END] = open;
        c[BSP_C_CROWDED_END] = crowded;
    }
    c[BSP_C_OPEN] = 0;
    c[BSP_C_CROWDED] = 0;
}

/* bookkeeping after a convergence check */
BSP_HD void bsp_check_ctl(const BspEigChunk &g, int iter)
{
    int *c = g.counters;
    if (c[BSP_C_REFINED]) return;
    c[BSP_C_ITERS] = iter + 1;
    c[BSP_C_UNCONV_END] = c[BSP_C_UNCONV];
    if (c[BSP_C_UNCONV] > c[BSP_C_SELECTED]) c[BSP_C_SELECTED] = c[BSP_C_UNCONV];
    if (c[BSP_C_UNCONV] == 0) c[BSP_C_REFINED] = 1;
    c[BSP_C_UNCONV] = 0;
    /* the list the check of iteration iter + 1 appends to starts empty */
    if (g.rcount) for (int p = 0; p < g.npencil; ++p) g.rcount[((iter + 1) & 1) * g.npencil + p] = 0;
}

/* ------------------------------------------------------------------------- *
 * Device-side state selection: which eigenvectors SOLVE_SYSTEM keeps (matrices.f90:296-334, KIND_PI >= 3 branch):
 *   n1_fin = #{E <= Emax_fin} + 1,  ntemp_raw = #{E <= Elim},  nlim = max over the l done so far of ntemp_raw,
 *   ntemp  = MIN(MAX(n1_fin + 40, nlim), nfun)        -> ctemp(:, 1:ntemp, l) = Hij(:, 1:ntemp)
 * With sel.mode = 1 the bracketing closes EVERY bracket to rounding first (nvec_br = 0), so the counts below are
 * taken on final eigenvalues and equal the reference's; the refinement then only computes columns 1..ntemp.
 * One thread walks the pencils of the chunk in the caller's order (the running maximum nlim is sequential over
 * l; pencils of one selection group are kept in one chunk).
 * ------------------------------------------------------------------------- */
struct BspSelect {
    int mode;      /* 0: keep the caller's nvec */
    int extra;     /* 41 for the reference's "n1_fin + 40" (n1_fin itself is the count + 1) */
    int group;     /* running-maximum group (>= 0) or -1 */
    int cap;       /* the caller's nvec: upper limit */
    double ecut_a; /* Emax_fin */
    double ecut_b; /* Elim     */
};

BSP_HD int bsp_count_below(const double *lo, const double *hi, int n, double cut)
{
    /* #{e: E_e <= cut}, E_e = bracket midpoint, ascending */
    int a = -1, b = n;
    while (b - a > 1) {
        const int m = (a + b) >> 1;
        if (0.5 * (lo[m] + hi[m]) <= cut) a = m; else b = m;
    }
    return a + 1;
}

BSP_HD void bsp_select_states(const BspEigChunk &g, const BspSelect *sel, int *nvec_out, int *nsel_report)
{
    const int buf = g.counters[BSP_C_BUF];
    const size_t per = (size_t)g.npencil * g.ldw;
    int cur_group = -2, runmax = 0;
    for (int p = 0; p < g.npencil; ++p) {
        const BspSelect s = sel[p];
        int nv = s.cap;
        if (s.mode == 1) {
            const double *lo = g.lo + (size_t)buf * per + (size_t)p * g.ldw, *hi = g.hi + (size_t)buf * per + (size_t)p * g.ldw;
            const int ca = bsp_count_below(lo, hi, g.n, s.ecut_a), cb = bsp_count_below(lo, hi, g.n, s.ecut_b);
            if (s.group != cur_group) { cur_group = s.group; runmax = 0; }
            if (cb > runmax) runmax = cb;
            int want = ca + s.extra;
            if (s.group >= 0 && runmax > want) want = runmax;
            else if (s.group < 0 && cb > want) want = cb;
            nv = want < s.cap ? want : s.cap;
            if (nv < 0) nv = 0;
        }
        nvec_out[p] = nv;
        if (nsel_report) nsel_report[p] = nv;
    }
}

/* ------------------------------------------------------------------------- *
 * hand-over: final brackets live in buffer `buf`; copy to buffer 0, set the
 * first shift to the bracket midpoint.
 * ------------------------------------------------------------------------- */
BSP_HD void bsp_refine_prepare(const BspEigChunk &g, int p, int e)
{
    if (e >= g.n) return;
    const int buf = g.counters[BSP_C_BUF];
    const size_t per = (size_t)g.npencil * g.ldw;
    const size_t id = (size_t)p * g.ldw + e;
    const double lo = g.lo[(size_t)buf * per + id], hi = g.hi[(size_t)buf * per + id];
    if (buf != 0) { g.lo[id] = lo; g.hi[id] = hi; }
    const double mid = 0.5 * (lo + hi);
    g.sigma[id] = mid;
    g.rho[id] = mid;
    g.rho_prev[id] = mid;
    g.scale[id] = 1.0;
    g.res[id] = INFINITY;
    g.status[id] = (e < g.nvec[p]) ? 0 : BSP_ST_CONVERGED;
}

/* ------------------------------------------------------------------------- *
 * F pass: LDL^T of H - sigma S fused with the forward substitution of the
 * right-hand side; stores (zd, l_1..l_B) per row.  Also refines the bracket
 * with the inertia it gets for free.
 *   iter == 0 : rhs = hashed uniform(-1,1)      (plain inverse iteration)
 *   iter  > 0 : rhs = scale * R  (R written by the previous B pass)
 * ------------------------------------------------------------------------- */
/* ls: column of the factor workspace L this thread uses (= e at full width, = its slot in a compacted pass).
 * CKPT: check-pointed form -- nothing is stored per row; at every segment start the state of the elimination
 * goes to CK[p][segment][0..BSP_CK_DOUBLES)[ls] (see BSP_CK_GROUPS). */
template <int B, bool CKPT = false, class Src>
BSP_HD void bsp_factor_forward_rows(const BspEigChunk &g, int p, int e, int ls, int iter, bool active, Src &src)
{
    constexpr int K1 = B + 1;
    constexpr int FS = 2 * B + 2;
    constexpr int TR = Src::TR;
    constexpr int CKD = BSP_CK_DOUBLES(B);
    static_assert(!CKPT || TR % BSP_CK_STEPS(B) == 0, "tiles hold whole check-point segments");
    const int n = g.n, npad = g.npad, ldw = g.ldw;
    const int ntiles = npad / TR;
    const size_t id = (size_t)p * g.ldw + (active ? e : 0);
    const double sigma = active ? g.sigma[id] : 0.0;
    const double sc = active ? g.scale[id] : 1.0;
    const double pivmin = 1e-30 * (g.pbound[p * 4 + 2] + fabs(sigma) * g.pbound[p * 4 + 3]);
    const double *__restrict__ Rp = g.R + (size_t)p * g.xrows * ldw + e;
    double *__restrict__ Lp = CKPT ? g.CK + (size_t)p * (npad / BSP_CK_STEPS(B)) * CKD * ldw + ls
                                   : g.L + (size_t)p * npad * K1 * ldw + ls;
    BSP_ASSERT(p >= 0 && p < g.npencil && (!active || (ls >= 0 && ls < ldw && e >= 0 && e < n)));   /* idle threads touch nothing */
    BSP_ASSERT(npad % TR == 0 && g.xrows >= npad && g.nrows >= npad + B + 1);

    double w[K1][K1], y[K1];
    int cnt = 0;
    /* raw right-hand side (unscaled): the multiply by `sc` happens when the value is consumed, so
     * that the load issued a whole block ahead has nothing waiting on it */
    const double scr = (iter == 0) ? 1.0 : sc;
    auto rhs = [&](int row) -> double {
        if (row >= n) return 0.0;
        BSP_ASSERT(row >= 0 && row < g.xrows);
        return (iter == 0) ? bsp_hash_uniform(0u, (uint32_t)e, (uint32_t)row) : Rp[(size_t)row * ldw];
    };
    /* software pipeline: right-hand side (HBM) a whole unrolled block (B+1 rows) ahead; the band row that
     * enters the window at the end of a step is read from the tile at the top of the step */
    double rq[K1];
    /* staged sources keep the right-hand side of the following groups in flight through a per-thread ring in
     * shared memory (asynchronous copies, no registers): the reads sit behind the kernel's own write stream in
     * the memory controller and take several microseconds, more than one group of look-ahead hides.  Only the
     * passes that read R use it (iteration 0 computes its start vector). */
    const bool ring = (Src::RHS_RING > 0) && iter > 0;
    auto ring_issue = [&](int gi) {
        if constexpr (Src::RHS_RING > 0) {
            const int slot = (gi % (Src::RHS_RING + 1)) * K1;
#pragma unroll
            for (int t = 0; t < K1; ++t) {
                const int row = gi * K1 + t;
                if (row < n) src.rhs_issue(slot + t, Rp + (size_t)row * ldw);
                else src.rhs_zero(slot + t);
            }
            src.rhs_commit();
        }
    };
    src.begin_forward(ntiles);
    for (int tl = 0; tl < ntiles; ++tl) {
        const double *tH, *tS;
        src.acquire_forward(tl, tH, tS);
        if (active) {
            if (tl == 0) {
                bsp_sturm_init<B, Src::GL>(w, tH, tS, sigma);
#pragma unroll
                for (int r = 0; r < K1; ++r) y[r] = scr * rhs(r);
                if constexpr (Src::RHS_RING > 0) {
                    /* groups 1 .. RHS_RING of the right-hand side go in flight through the shared-memory ring */
                    if (ring) {
                        for (int gi = 1; gi <= Src::RHS_RING; ++gi) ring_issue(gi);
                    } else {
#pragma unroll
                        for (int r = 0; r < K1; ++r) rq[r] = rhs(K1 + r);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < K1; ++r) rq[r] = rhs(K1 + r);
                }
            }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int gq = 0; gq < TR; gq += K1) {
                const int j0 = tl * TR + gq;
                const double *hnext = tH + (size_t)(gq + K1) * FS, *snext = tS + (size_t)(gq + K1) * FS;
                if constexpr (CKPT) {
                    /* with the shared-memory ring the values of the next group are fetched below, before they are
                     * needed by the steps: the check-point takes them from the ring as well */
                }
                if constexpr (Src::RHS_RING > 0) {
                    if (ring) {
                        /* this group consumes the rows of group gi+1; group gi+1+RHS_RING takes the slot that
                         * the previous group emptied */
                        const int gi = j0 / K1;
                        ring_issue(gi + 1 + Src::RHS_RING);
                        src.template rhs_wait<Src::RHS_RING>();
#pragma unroll
                        for (int t = 0; t < K1; ++t) rq[t] = src.rhs_read(((gi + 1) % (Src::RHS_RING + 1)) * K1 + t);
                    }
                }
                if constexpr (CKPT) {
                    if ((j0 / K1) % BSP_CK_GROUPS(B) == 0) {
                        /* at a group start the window slots are in identity position; row B of the window is
                         * still the plain band row (it entered with the last step) and is not stored */
                        double *ck = Lp + (size_t)(j0 / BSP_CK_STEPS(B)) * CKD * ldw;
                        BSP_ASSERT(g.CK != nullptr && j0 % BSP_CK_STEPS(B) == 0 && j0 / BSP_CK_STEPS(B) < npad / BSP_CK_STEPS(B));
                        int q = 0;
#pragma unroll
                        for (int r = 0; r < B; ++r) {
#pragma unroll
                            for (int c = 0; c <= r; ++c) { ck[(size_t)q * ldw] = w[r][c]; ++q; }
                        }
#pragma unroll
                        for (int r = 0; r < K1; ++r) { ck[(size_t)q * ldw] = y[r]; ++q; }
#pragma unroll
                        for (int r = 0; r < K1; ++r) { ck[(size_t)q * ldw] = rq[r]; ++q; }   /* raw rhs of rows j0+K1 .. j0+2K1-1 */
                    }
                }
#pragma unroll
                for (int t = 0; t < K1; ++t) {
                    const int j = j0 + t;
                    double nh[K1], ns[K1];
#pragma unroll
                    for (int m = 0; m <= B; ++m) {
                        nh[m] = bsp_ld<Src::GL>(hnext + (size_t)t * FS + m);
                        ns[m] = bsp_ld<Src::GL>(snext + (size_t)t * FS + m);
                    }
                    const double rnew = scr * rq[t];    /* rhs of row j+K1, loaded K1 steps ago */
                    if (!ring) rq[t] = rhs(j + 2 * K1);
                    double d = w[t][t];
                    if (fabs(d) < pivmin) d = -pivmin;
                    if (d < 0.0) ++cnt;
                    const double rinv = BSP_RCP(d);
                    double col[K1], l[K1];
                    const double y0 = y[t];
                    double *Lrow = Lp + (size_t)j * K1 * ldw;
                    BSP_ASSERT(j >= 0 && j < npad);
                    if (!CKPT) Lrow[0] = y0 * rinv;
#pragma unroll
                    for (int i = 1; i <= B; ++i) {
                        col[i] = w[(t + i) % K1][t];
                        l[i] = col[i] * rinv;
                        if (!CKPT) Lrow[(size_t)i * ldw] = l[i];
                        y[(t + i) % K1] = fma(-l[i], y0, y[(t + i) % K1]);
                    }
#pragma unroll
                    for (int m = 1; m <= B; ++m) {
#pragma unroll
                        for (int i = m; i <= B; ++i) {
                            w[(t + i) % K1][(t + m) % K1] = fma(-l[i], col[m], w[(t + i) % K1][(t + m) % K1]);
                        }
                    }
#pragma unroll
                    for (int m = 0; m <= B; ++m) w[t][(t + 1 + m) % K1] = fma(-sigma, ns[m], nh[m]);
                    y[t] = rnew;
                }
            }
        }
        src.release_forward(tl, ntiles);
    }
    /* inertia -> bracket (buffer 0) */
    if (active) {
        if (cnt <= e) { if (sigma > g.lo[id]) g.lo[id] = sigma; }
        else { if (sigma < g.hi[id]) g.hi[id] = sigma; }
    }
}

BSP_HD bool bsp_refine_active(const BspEigChunk &g, int p, int e)
{
    return e < g.n && !(g.status[(size_t)p * g.ldw + e] & BSP_ST_CONVERGED);
}

template <int B>
BSP_HD void bsp_factor_forward(const BspEigChunk &g, int p, int e, int iter)
{
    if (!bsp_refine_active(g, p, e)) return;
    BspRowsGlobal<B> src{g.fbH + (size_t)p * g.nrows * (2 * B + 2), g.fbS + (size_t)g.inst[p] * g.nrows * (2 * B + 2)};
    bsp_factor_forward_rows<B>(g, p, e, e, iter, true, src);
}

/* ------------------------------------------------------------------------- *
 * B pass: back substitution  y = L^-T (zd);  x_new = cx*scale*x_old - y;
 * banded matvecs s = S x_new, h = H x_new on a sliding window; Rayleigh
 * quotient, residual, next right-hand side and next shift.
 *   corr_now  : this iteration is in correction form (cx = 1) else plain (cx = 0)
 *   corr_next : what to leave in R for the next F pass:
 *               1 -> h - rho' s  (rho' = Rayleigh quotient known at pass start)
 *               0 -> s
 * ------------------------------------------------------------------------- */
/* end of a solving back sweep: Rayleigh quotient, normalisation, residual norms, next shift */
/* corr_next < 0 (the check follows at once): the sweep only knows r = (H - rho' S) x against the PREVIOUS quotient rho';
 * the residual against the new one is r - (rho - rho') S x, whose 2-norm follows from the three sums the sweep carried,
 *   ||r_new||^2 = rr2 - 2 (rho - rho') rs + (rho - rho')^2 s2,    rs = sum r_i (Sx)_i,  s2 = sum (Sx)_i^2,
 * so the convergence / selection test needs no residual pass over all eigenpairs (round 2 until this change: a
 * latency-bound pass of 3 ms per 408 pencils).  The rounding of the three sums is added, so the figure is an upper bound:
 * a pair near the threshold is selected for the correction pass rather than missed; res (max norm) <= res2. */
BSP_HD void bsp_back_finish(const BspEigChunk &g, size_t id, int corr_next, double sc, double rho_p, double xSx, double xHx,
                            double resmax, double rr2, double xabs, double rs, double s2)
{
    double lo = g.lo[id], hi = g.hi[id];
    const double good = (xSx > 0.0 && xSx < INFINITY) ? 1.0 : 0.0;
    double rho_new = rho_p, scn = sc, res = INFINITY, res2 = INFINITY;
    if (good != 0.0) {
        rho_new = xHx / xSx;
        scn = 1.0 / sqrt(xSx);
        res = resmax * scn;
        res2 = sqrt(rr2) * scn;
        if (corr_next < 0) {
            const double dr = rho_new - rho_p;
            const double t1 = 2.0 * dr * rs, t2 = dr * dr * s2;
            const double q = fmax(rr2 - t1 + t2, 0.0) + 8.0 * BSP_EPS * (rr2 + fabs(t1) + t2);
            res2 = sqrt(q) * scn;
            res = res2;
        }
    }
    g.xmax[id] = xabs;
    g.rho_prev[id] = rho_p;
    g.rho[id] = rho_new;
    g.scale[id] = scn;
    g.res[id] = res;
    g.res2[id] = res2;
    double sig = (rho_new > lo && rho_new < hi) ? rho_new : 0.5 * (lo + hi);
    if (corr_next > 0) {
        double gp = g.gap[id];
        if (!(gp > 0.0) || !(gp < INFINITY)) gp = fmax(hi - lo, fabs(rho_new) * 1e-6 + 1e-12);
        const double delta = g.delta_rel * gp;
        if (fabs(sig - rho_p) < delta) sig = (rho_p + delta < hi) ? rho_p + delta : rho_p - delta;
    }
    g.sigma[id] = sig;
}

#ifndef BSP_BACK_PF
#define BSP_BACK_PF 2 /* factor rows in flight per thread in the back sweep */
#endif

/* RESID = true turns the pass into a pure residual evaluation of the vector already in X: nothing is solved and
 * X is not written; r = (H - rho S) x with the CURRENT Rayleigh quotient goes to R (ready for a correction
 * step) and its scaled max norm to res.  The plain passes only know the residual against the previous
 * quotient, so this cheap pass (no factor traffic) is what lets the iteration stop after the second solve. */
template <int B, bool RESID = false, class Src>
BSP_HD void bsp_back_substitute_rows(const BspEigChunk &g, int p, int e, int ls, int corr_now, int corr_next, bool active, Src &src)
{
    constexpr int K1 = B + 1;
    constexpr int FS = 2 * B + 2;
    constexpr int PF = RESID ? 2 * K1 : BSP_BACK_PF; /* rows in flight per thread */
    constexpr int TR = Src::TR;
    static_assert(TR % PF == 0, "tiles hold whole groups of PF steps");
    const size_t id = (size_t)p * g.ldw + (active ? e : 0);
    const int n = g.n, npad = g.npad, ldw = g.ldw;
    const int ntiles = npad / TR;
    const double *__restrict__ Lp = g.L + (size_t)p * npad * K1 * ldw + ls;
    double *__restrict__ Xp = g.X + (size_t)p * g.xrows * ldw + e;
    double *__restrict__ Rp = g.R + (size_t)p * g.xrows * ldw + e;
    BSP_ASSERT(p >= 0 && p < g.npencil && (!active || (ls >= 0 && ls < ldw && e >= 0 && e < n)) && npad % TR == 0);
    const double sc = active ? g.scale[id] : 1.0;
    const double rho_p = active ? g.rho[id] : 0.0; /* rho' */
    const double cx = corr_now ? sc : 0.0;

    /* windows over rows j .. j+B (index 0 = row j after the shift of step j):
     *   yw : solution of L^T y = zd         xv : x_new
     *   hs, ss : partial sums of (H x_new)_i and (S x_new)_i.
     * Column sweep: when x_new[j] appears, column j of the (symmetric) band adds A(i,j) x_j to the
     * rows i = j+1..j+B and row j collects its upper part sum_d A(j,j+d) x_{j+d}; row i is complete
     * once x_{i-B} is in, i.e. at step j = i-B.  Only the B+1 entries A(j, j..j+B) of each matrix
     * are read per step (half of a full-band row). */
    double yw[K1], xv[K1], hs[K1], ss[K1];
#pragma unroll
    for (int i = 0; i < K1; ++i) { yw[i] = 0.0; xv[i] = 0.0; hs[i] = 0.0; ss[i] = 0.0; }
    double xSx = 0.0, xHx = 0.0, resmax = 0.0, xabs = 0.0, rr2 = 0.0, rs = 0.0, s2 = 0.0;

    /* ring of PF factor rows (+ x_old) in flight: row j is consumed PF steps after its loads were
     * issued, which is what hides the HBM latency of this purely streaming sweep */
    double Lq[PF][K1], xq[PF];
    auto fetch = [&](int q, int row) {
        /* rows -PF..-1 are asked for by the last steps and never used: row 0 is loaded again instead */
        const int r = row < 0 ? 0 : row;
        BSP_ASSERT(r < npad);
        if (!RESID) {
            const double *Lrow = Lp + (size_t)r * K1 * ldw;
#pragma unroll
            for (int i = 0; i <= B; ++i) Lq[q][i] = Lrow[(size_t)i * ldw];
        }
        xq[q] = ((RESID || corr_now) && r < n) ? Xp[(size_t)r * ldw] : 0.0;
    };
    if (active) {
#pragma unroll
        for (int q = 0; q < PF; ++q) fetch(q, npad - 1 - q);
    }
    /* one step; rowH/rowS = band row j (its upper half A(j, j..j+B) is column j of the symmetric band).
     * The row is read at the top of the step from the staged tile: the shared-memory latency hides behind the
     * dependent chain of the substitution, and nothing is carried in registers across steps.
     * TAIL = false: a step of the tiles (0 <= j < npad);
     * TAIL = true: one of the last B steps (j < 0), which only drain the windows. */
    auto step = [&](auto tail_c, int j, int q, const double *rowH, const double *rowS) {
        constexpr bool TAIL = decltype(tail_c)::value;
        double ah[K1], as[K1];
        if (!TAIL) {
#pragma unroll
            for (int d = 0; d <= B; ++d) {
                ah[d] = bsp_ld<Src::GL>(rowH + B + d);
                as[d] = bsp_ld<Src::GL>(rowS + B + d);
            }
        }
        /* shift the windows: index 0 becomes row j */
#pragma unroll
        for (int i = B; i >= 1; --i) { yw[i] = yw[i - 1]; xv[i] = xv[i - 1]; hs[i] = hs[i - 1]; ss[i] = ss[i - 1]; }
        double xn = 0.0;
        if (!TAIL) {
            if (RESID) {
                xn = xq[q];
            } else {
                double yj = Lq[q][0];
#pragma unroll
                for (int i = B; i >= 1; --i) yj = fma(-Lq[q][i], yw[i], yj);   /* yw[i] = y_{j+i} */
                yw[0] = yj;
                if (j < n) {
                    xn = fma(cx, xq[q], -yj);
                    BSP_ASSERT(j >= 0 && j < g.xrows);
                    Xp[(size_t)j * ldw] = xn;
                    xabs = fmax(xabs, fabs(xn));
                }
            }
            fetch(q, j - PF);   /* slot q is free again: row j-PF goes in flight */
        } else {
            yw[0] = 0.0;
        }
        xv[0] = xn;
        /* column j: A(j+d, j) = A(j, j+d) = ah[d] */
        double h0 = 0.0, s0 = 0.0;
        if (!TAIL) {
#pragma unroll
            for (int d = 1; d <= B; ++d) {
                hs[d] = fma(ah[d], xn, hs[d]);
                ss[d] = fma(as[d], xn, ss[d]);
            }
#pragma unroll
            for (int d = 0; d <= B; ++d) {
                h0 = fma(ah[d], xv[d], h0);
                s0 = fma(as[d], xv[d], s0);
            }
        }
        hs[0] = h0;
        ss[0] = s0;
        /* row i = j + B is complete */
        const int i = j + B;
        if (i < n) {
            const double h = hs[B], sv = ss[B], xi = xv[B];
            xSx = fma(xi, sv, xSx);
            xHx = fma(xi, h, xHx);
            const double r = fma(-rho_p, sv, h);
            resmax = fmax(resmax, fabs(r));
            rr2 = fma(r, r, rr2);
            rs = fma(r, sv, rs);
            s2 = fma(sv, sv, s2);
            /* corr_next < 0: a residual pass follows and writes R itself */
            if (RESID || corr_next >= 0) Rp[(size_t)i * ldw] = (RESID || corr_next) ? r : sv;
        }
    };
    src.begin_backward(ntiles);
    for (int tl = ntiles - 1; tl >= 0; --tl) {
        const double *tH, *tS;   /* band row tl*TR */
        src.acquire_backward(tl, ntiles, tH, tS);
        if (active) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int jl = TR - 1; jl >= 0; jl -= PF) {
#pragma unroll
                for (int q = 0; q < PF; ++q) {
                    const int off = (jl - q) * FS;
                    step(BspFalse(), tl * TR + jl - q, q, tH + off, tS + off);
                }
            }
        }
        src.release_backward(tl, ntiles);
    }
    if (!active) return;
    /* the last B steps only drain the windows */
    for (int j0 = -1; j0 >= -B; j0 -= PF) {
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            const int j = j0 - q;
            if (j >= -B) step(BspTrue(), j, q, nullptr, nullptr);
        }
    }
    if (RESID) {
        const bool ok = (xSx > 0.0 && xSx < INFINITY);
        g.res[id] = ok ? resmax * sc : INFINITY;
        g.res2[id] = ok ? sqrt(rr2) * sc : INFINITY;
        /* the correction pass that may follow subtracts M^-1 (H - rho S) x with THIS rho: its shift must keep
         * the distance delta from it (same rule as at the end of a solving pass) */
        double lo = g.lo[id], hi = g.hi[id], sig = g.sigma[id];
        double gp = g.gap[id];
        if (!(gp > 0.0) || !(gp < INFINITY)) gp = fmax(hi - lo, fabs(rho_p) * 1e-6 + 1e-12);
        const double delta = g.delta_rel * gp;
        if (fabs(sig - rho_p) < delta) sig = (rho_p + delta < hi) ? rho_p + delta : rho_p - delta;
        g.sigma[id] = sig;
        return;
    }
    bsp_back_finish(g, id, corr_next, sc, rho_p, xSx, xHx, resmax, rr2, xabs, rs, s2);
}

template <int B>
BSP_HD void bsp_back_substitute(const BspEigChunk &g, int p, int e, int corr_now, int corr_next)
{
    if (!bsp_refine_active(g, p, e)) return;
    BspRowsGlobal<B> src{g.fbH + (size_t)p * g.nrows * (2 * B + 2), g.fbS + (size_t)g.inst[p] * g.nrows * (2 * B + 2)};
    bsp_back_substitute_rows<B>(g, p, e, e, corr_now, corr_next, true, src);
}

template <int B>
BSP_HD void bsp_residual_pass(const BspEigChunk &g, int p, int e)
{
    if (!bsp_refine_active(g, p, e)) return;
    BspRowsGlobal<B> src{g.fbH + (size_t)p * g.nrows * (2 * B + 2), g.fbS + (size_t)g.inst[p] * g.nrows * (2 * B + 2)};
    bsp_back_substitute_rows<B, true>(g, p, e, e, 0, 1, true, src);
}

/* ------------------------------------------------------------------------- *
 * Check-pointed B pass (plain solves only: corr_now = 0).  Segments of BSP_CK_STEPS(B) rows, last to first:
 *   phase A  restore the elimination state from the check-point the forward sweep left (rows 0..B-1 of the pivot
 *            window, y[0..B], the raw rhs of the following group; row B of the window is re-read from the band) and
 *            re-eliminate the segment: (zd, l_1..l_B) of its rows go to the scratch;
 *   phase B  back substitution + column-sweep band matvecs over the segment, exactly the steps of
 *            bsp_back_substitute_rows with the factor rows taken from the scratch.
 * Same arithmetic in the same order as the stored-factor pair of sweeps: bit-identical results.
 * Scratch: `Scr` hands out slot(i), i < BSP_CK_STEPS*(B+1) -- a thread-private array on the host replay, a column of
 * shared memory in the kernel, where the check-point of the NEXT segment is also prefetched (cp.async) into the
 * slots phase B has already consumed.
 * ------------------------------------------------------------------------- */
template <int B>
struct BspScratchLocal {
    static constexpr int SLOTS = BSP_CK_STEPS(B) * (B + 1);
    double s[SLOTS];
    BSP_HD double &slot(int i) { BSP_ASSERT(i >= 0 && i < SLOTS); return s[i]; }
    /* check-point values are read straight from global memory: no prefetch on the host */
    BSP_HD void prefetch(int, const double *) {}
    BSP_HD void prefetch_commit() {}
    BSP_HD void prefetch_wait() {}
    BSP_HD double fetched(int, const double *gp) { return *gp; }
};

template <int B, class Src, class Scr>
BSP_HD void bsp_back_ckpt_rows(const BspEigChunk &g, int p, int e, int ls, int iter, int corr_next, bool active, Src &src, Scr &scr)
{
    constexpr int K1 = B + 1;
    constexpr int FS = 2 * B + 2;
    constexpr int SG = BSP_CK_GROUPS(B);
    constexpr int ST = BSP_CK_STEPS(B);
    constexpr int CKD = BSP_CK_DOUBLES(B);
    constexpr int TR = Src::TR;
    static_assert(TR == ST, "the check-pointed back sweep takes one segment per tile");
    constexpr int PRE_ROWS = (CKD + K1 - 1) / K1;   /* scratch rows (from the top) that receive the next check-point */
    static_assert(PRE_ROWS < ST, "segment too short to prefetch a check-point into its own scratch");
    const size_t id = (size_t)p * g.ldw + (active ? e : 0);
    const int n = g.n, npad = g.npad, ldw = g.ldw;
    const int nseg = npad / ST;
    const double *__restrict__ Cp = g.CK + (size_t)p * nseg * CKD * ldw + ls;
    double *__restrict__ Xp = g.X + (size_t)p * g.xrows * ldw + e;
    double *__restrict__ Rp = g.R + (size_t)p * g.xrows * ldw + e;
    const double sc = active ? g.scale[id] : 1.0;
    const double rho_p = active ? g.rho[id] : 0.0;
    const double sigma = active ? g.sigma[id] : 0.0;
    const double pivmin = 1e-30 * (g.pbound[p * 4 + 2] + fabs(sigma) * g.pbound[p * 4 + 3]);
    /* the forward sweep scaled the raw right-hand side when it consumed it: iteration 0 (hashed start vector) by 1,
     * later ones by the normalisation of the previous vector.  rho' and scale are untouched between the two sweeps. */
    const double scr_f = (iter == 0) ? 1.0 : sc;

    double yw[K1], xv[K1], hs[K1], ss[K1];
#pragma unroll
    for (int i = 0; i < K1; ++i) { yw[i] = 0.0; xv[i] = 0.0; hs[i] = 0.0; ss[i] = 0.0; }
    double xSx = 0.0, xHx = 0.0, resmax = 0.0, xabs = 0.0, rr2 = 0.0, rs = 0.0, s2 = 0.0;

    auto back_step = [&](auto tail_c, int j, const double *rowH, const double *rowS, int srow) {
        constexpr bool TAIL = decltype(tail_c)::value;
        double ah[K1], as[K1];
        if (!TAIL) {
#pragma unroll
            for (int d = 0; d <= B; ++d) {
                ah[d] = bsp_ld<Src::GL>(rowH + B + d);
                as[d] = bsp_ld<Src::GL>(rowS + B + d);
            }
        }
#pragma unroll
        for (int i = B; i >= 1; --i) { yw[i] = yw[i - 1]; xv[i] = xv[i - 1]; hs[i] = hs[i - 1]; ss[i] = ss[i - 1]; }
        double xn = 0.0;
        if (!TAIL) {
            double yj = scr.slot(srow * K1);
#pragma unroll
            for (int i = B; i >= 1; --i) yj = fma(-scr.slot(srow * K1 + i), yw[i], yj);
            yw[0] = yj;
            if (j < n) {
                xn = 0.0 - yj;   /* = fma(0, x_old, -y) of the stored-factor sweep, bit for bit */
                Xp[(size_t)j * ldw] = xn;
                xabs = fmax(xabs, fabs(xn));
            }
        } else {
            yw[0] = 0.0;
        }
        xv[0] = xn;
        double h0 = 0.0, s0 = 0.0;
        if (!TAIL) {
#pragma unroll
            for (int d = 1; d <= B; ++d) {
                hs[d] = fma(ah[d], xn, hs[d]);
                ss[d] = fma(as[d], xn, ss[d]);
            }
#pragma unroll
            for (int d = 0; d <= B; ++d) {
                h0 = fma(ah[d], xv[d], h0);
                s0 = fma(as[d], xv[d], s0);
            }
        }
        hs[0] = h0;
        ss[0] = s0;
        const int i = j + B;
        if (i < n) {
            const double h = hs[B], sv = ss[B], xi = xv[B];
            xSx = fma(xi, sv, xSx);
            xHx = fma(xi, h, xHx);
            const double r = fma(-rho_p, sv, h);
            resmax = fmax(resmax, fabs(r));
            rr2 = fma(r, r, rr2);
            rs = fma(r, sv, rs);
            s2 = fma(sv, sv, s2);
            if (corr_next >= 0) Rp[(size_t)i * ldw] = corr_next ? r : sv;
        }
    };

    src.begin_backward_wide(nseg);
    /* check-point of the last segment: nothing to hide its latency behind */
    if (active) {
        const double *ck = Cp + (size_t)(nseg - 1) * CKD * ldw;
        for (int q = 0; q < CKD; ++q) scr.prefetch((ST - PRE_ROWS) * K1 + q, ck + (size_t)q * ldw);
        scr.prefetch_commit();
    }
    for (int seg = nseg - 1; seg >= 0; --seg) {
        const double *tH, *tS;   /* band rows seg*ST .. seg*ST + ST + B */
        src.acquire_backward_wide(seg, nseg, tH, tS);
        if (active) {
            const int js = seg * ST;
            /* ---- phase A: re-eliminate rows js .. js+ST-1 ---- */
            double w[K1][K1], y[K1], rq[K1];
            {
                const double *ck = Cp + (size_t)seg * CKD * ldw;
                scr.prefetch_wait();
                int q = 0;
#pragma unroll
                for (int r = 0; r < B; ++r) {
#pragma unroll
                    for (int c = 0; c < K1; ++c) {
                        if (c <= r) { w[r][c] = scr.fetched((ST - PRE_ROWS) * K1 + q, ck + (size_t)q * ldw); ++q; } else w[r][c] = 0.0;
                    }
                }
#pragma unroll
                for (int r = 0; r < K1; ++r) { y[r] = scr.fetched((ST - PRE_ROWS) * K1 + q, ck + (size_t)q * ldw); ++q; }
#pragma unroll
                for (int r = 0; r < K1; ++r) { rq[r] = scr.fetched((ST - PRE_ROWS) * K1 + q, ck + (size_t)q * ldw); ++q; }
                /* row B of the window: the plain band row js + B */
#pragma unroll
                for (int m = 0; m <= B; ++m)
                    w[B][m] = fma(-sigma, bsp_ld<Src::GL>(tS + (size_t)B * FS + m), bsp_ld<Src::GL>(tH + (size_t)B * FS + m));
            }
#pragma unroll
            for (int gq = 0; gq < SG; ++gq) {
#pragma unroll
                for (int t = 0; t < K1; ++t) {
                    const int jj = gq * K1 + t;
                    double nh[K1], ns[K1];
#pragma unroll
                    for (int m = 0; m <= B; ++m) {
                        nh[m] = bsp_ld<Src::GL>(tH + (size_t)(jj + K1) * FS + m);
                        ns[m] = bsp_ld<Src::GL>(tS + (size_t)(jj + K1) * FS + m);
                    }
                    /* the right-hand side of row js+jj+K1 matters only while that row belongs to this segment */
                    const double rnew = (gq + 1 < SG) ? scr_f * rq[t] : 0.0;
                    double d = w[t][t];
                    if (fabs(d) < pivmin) d = -pivmin;
                    const double rinv = BSP_RCP(d);
                    double col[K1], l[K1];
                    const double y0 = y[t];
                    scr.slot(jj * K1) = y0 * rinv;
#pragma unroll
                    for (int i = 1; i <= B; ++i) {
                        col[i] = w[(t + i) % K1][t];
                        l[i] = col[i] * rinv;
                        scr.slot(jj * K1 + i) = l[i];
                        y[(t + i) % K1] = fma(-l[i], y0, y[(t + i) % K1]);
                    }
#pragma unroll
                    for (int m = 1; m <= B; ++m) {
#pragma unroll
                        for (int i = m; i <= B; ++i) {
                            w[(t + i) % K1][(t + m) % K1] = fma(-l[i], col[m], w[(t + i) % K1][(t + m) % K1]);
                        }
                    }
#pragma unroll
                    for (int m = 0; m <= B; ++m) w[t][(t + 1 + m) % K1] = fma(-sigma, ns[m], nh[m]);
                    y[t] = rnew;
                }
            }
            /* ---- phase B: back substitution + matvecs, last row of the segment first ---- */
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int jj = ST - 1; jj >= 0; --jj) {
                back_step(BspFalse(), js + jj, tH + (size_t)jj * FS, tS + (size_t)jj * FS, jj);
                if (jj == ST - PRE_ROWS && seg > 0) {
                    /* the top PRE_ROWS scratch rows are consumed: the next check-point goes in flight into them */
                    const double *ck = Cp + (size_t)(seg - 1) * CKD * ldw;
                    for (int q = 0; q < CKD; ++q) scr.prefetch((ST - PRE_ROWS) * K1 + q, ck + (size_t)q * ldw);
                    scr.prefetch_commit();
                }
            }
        }
        src.release_backward_wide(seg, nseg);
    }
    if (!active) return;
    for (int j = -1; j >= -B; --j) back_step(BspTrue(), j, nullptr, nullptr, 0);
    bsp_back_finish(g, id, corr_next, sc, rho_p, xSx, xHx, resmax, rr2, xabs, rs, s2);
}

/* convergence bookkeeping after a B / residual pass (separate tiny kernel): marks eigenpairs whose scaled residual
 * is below conv_tol.  select != 0 (the check that follows the second solve + residual pass) additionally keeps an
 * eigenpair in the iteration when ||r||_2 / gap > vec_tol: the un-pivoted LDL^T leaves a few per cent of the
 * vectors with an error ~ ||r|| ||S^-1||^1/2 / gap of 1e-10 .. 1e-8 after two solves (they spoil the S-orthogonality of
 * their neighbours); exactly those get the correction pass that everybody got in round 1.  Measured on the
 * N = 1000 pencils (CPU replay): vec_tol = 1e-12 selects 10-17 % and |C^T S C - I| stays below 1e-11.
 * Unconverged eigen indices are appended to the compaction list the next pass runs on (order varies from run
 * to run; the result of an eigen index does not depend on its slot). */
/* the test itself: true when eigen index e stays in the iteration (marks it converged otherwise) */
BSP_HD bool bsp_check_keep(const BspEigChunk &g, int p, int e, int select)
{
    if (e >= g.n) return false;
    const size_t id = (size_t)p * g.ldw + e;
    if (g.status[id] & BSP_ST_CONVERGED) return false;
    const double rho = g.rho[id];
    const double r = g.res[id], a = fmax(1.0, fabs(rho));
    bool ok = (r <= g.conv_tol * a);
    if (ok && select) {
        /* gap to the neighbouring eigenvalues: their Rayleigh quotients where they are being refined, else the
         * distance of the brackets at hand-over */
        double gp = INFINITY;
        const int nv = g.nvec[p];
        if (e > 0 && e - 1 < nv) gp = fmin(gp, fabs(rho - g.rho[id - 1]));
        if (e + 1 < g.n && e + 1 < nv) gp = fmin(gp, fabs(g.rho[id + 1] - rho));
        const double gb = g.gap[id];
        if (gb > 0.0 && gb < gp && (e == 0 || e + 1 >= nv)) gp = gb;
        if (!(gp > 0.0)) gp = 0.0;
        /* select = 1: after the second solve; select = 2: after a correction pass -- a vector whose correction has not
         * brought ||r||_2 / gap within 100 vec_tol yet gets another one (the un-pivoted factor of a near-threshold level
         * can have pivot growth ~1e7: one correction then gains only ~1e-3 and leaves |c_i^T S c_j| ~ 1e-9, found on one
         * of the 4096 cfg3 problems); pairs that rounding cannot separate (gp ~ 0) are left to the degenerate-pair flag */
        if (select == 1) ok = (g.res2[id] <= g.vec_tol * gp);
        else if (gp > 64.0 * BSP_EPS * fmax(fabs(rho), 1e-300))
            /* ... unless the residual already sits at its rounding floor (1e-2 conv_tol ~ 500 eps): vectors of dense
             * spectra (cfg4: gaps ~1e-6) are there after one correction and more passes only stir the noise */
            ok = (g.res2[id] <= 100.0 * g.vec_tol * gp) || (g.res2[id] <= 1e-2 * g.conv_tol * a);
    }
    if (ok) g.status[id] |= BSP_ST_CONVERGED;
    return !ok;
}

/* host replay form: test + append to the compaction list (the kernel appends warp by warp, bsp_check_kernel) */
BSP_HD void bsp_check_converged(const BspEigChunk &g, int p, int e, int iter, int select)
{
    if (!bsp_check_keep(g, p, e, select)) return;
#if defined(__CUDA_ARCH__)
    atomicAdd(g.counters + BSP_C_UNCONV, 1);
    if (g.rlist) {
        const int slot = atomicAdd(g.rcount + (iter & 1) * g.npencil + p, 1);
        BSP_ASSERT(slot >= 0 && slot < g.n);
        g.rlist[((size_t)(iter & 1) * g.npencil + p) * g.ldw + slot] = e;
    }
#else
    g.counters[BSP_C_UNCONV] += 1;
    if (g.rlist) g.rlist[((size_t)(iter & 1) * g.npencil + p) * g.ldw + g.rcount[(iter & 1) * g.npencil + p]++] = e;
#endif
}

/* a compacted pass of iteration `iter` maps thread `slot` of pencil p to the eigen index the check of
 * iteration iter - 1 listed; returns -1 beyond the list */
BSP_HD int bsp_listed_index(const BspEigChunk &g, int p, int slot, int iter)
{
    const int buf = (iter - 1) & 1;
    if (slot >= g.rcount[buf * g.npencil + p]) return -1;
    BSP_ASSERT(g.rcount[buf * g.npencil + p] <= g.n);
    return g.rlist[((size_t)buf * g.npencil + p) * g.ldw + slot];
}

/* ------------------------------------------------------------------------- *
 * finalize, per eigenpair: eigenvalue, normalisation factor with the sign
 * convention (first coefficient with |c_i| >= 1e-6 max|c| positive).
 * fac[id] multiplies column e of X when it is transposed into C.
 * ------------------------------------------------------------------------- */
BSP_HD void bsp_finalize_eigen(const BspEigChunk &g, int p, int e, double *E, double *fac, int *bad,
                               double res_tol)
{
    if (e >= g.n) return;
    const size_t id = (size_t)p * g.ldw + e;
    if (e >= g.nvec[p]) {
        /* values only: the midpoint of a bracket the bracketing closed to 4 eps; anything wider is reported */
        const double lo = g.lo[id], hi = g.hi[id];
        E[(size_t)p * g.n + e] = 0.5 * (lo + hi);
        fac[id] = 0.0;
        if (!(hi - lo <= fmax(16.0 * BSP_EPS * fmax(fabs(lo), fabs(hi)), 4.0 * BSP_ABS_CLOSE))) {
#if defined(__CUDA_ARCH__)
            atomicAdd(bad + p, 1);
#else
            bad[p] += 1;
#endif
        }
        return;
    }
    const double rho = g.rho[id];
    E[(size_t)p * g.n + e] = rho;
    const double *Xp = g.X + (size_t)p * g.xrows * g.ldw + e;
    const double amax = g.xmax[id];   /* = max_j |X(j,e)|, from the back sweep that wrote the vector */
    double sgn = 1.0;
    const double thr = 1e-6 * amax;
    for (int j = 0; j < g.n; ++j) {
        const double v = Xp[(size_t)j * g.ldw];
        if (fabs(v) >= thr) { sgn = (v < 0.0) ? -1.0 : 1.0; break; }
    }
    fac[id] = sgn * g.scale[id];
    const double r = g.res[id];
    const bool unbracketed = g.pbound[p * 4 + 0] > g.pbound[p * 4 + 1];
    /* the Rayleigh quotient must sit in the bracket the inertia counts certify for index e: a vector that drifted
     * to a neighbouring eigenpair is reported, not returned.  Counts taken within ~100 ulp of an eigenvalue round
     * either way (pivot growth), so the bracket is widened by a tenth of the gap to the neighbouring brackets. */
    double gpf = g.gap[id];
    if (!(gpf > 0.0) || !(gpf < INFINITY)) gpf = 0.0;
    const double slack = fmax(0.1 * gpf, 4.0 * r + 1024.0 * BSP_EPS * fmax(fabs(g.lo[id]), fabs(g.hi[id]))) + 1e-300;
    const bool outside = !(rho >= g.lo[id] - slack && rho <= g.hi[id] + slack);
#if defined(BSP_TRACE_FIN) && !defined(__CUDA_ARCH__)
    if (outside) printf("outside: e=%d rho=%.17g lo=%.17g hi=%.17g r=%.3e slack=%.3e viol=%.3e\n", e, rho, g.lo[id], g.hi[id], r, slack, fmax(g.lo[id]-rho, rho-g.hi[id]));
#endif
    /* every vector comes from its own inverse iteration: nothing re-orthogonalises a numerically degenerate pair
     * (gap below a few hundred ulps), so such a pair is reported instead of returned as if it were S-orthogonal
     * (radial Sturm-Liouville spectra are simple; the LAPACK-shaped entry accepts any banded pencil) */
    bool degenerate = false;
    if (e + 1 < g.nvec[p] && e + 1 < g.n) degenerate = fabs(g.rho[id + 1] - rho) <= 256.0 * BSP_EPS * fmax(fabs(rho), fabs(g.rho[id + 1]));
    if (unbracketed || outside || degenerate || !(r <= res_tol * fmax(1.0, fabs(rho)))) {
#if defined(__CUDA_ARCH__)
        atomicAdd(bad + p, 1);
#else
        bad[p] += 1;
#endif
    }
}

#endif /* BSP_CORE_H */
