/*
 * bsp_driver.h -- stage schedule of the banded eigensolver for one chunk of
 * pencils, written against an executor so that the CUDA launcher
 * (bsp_api.cu) and the CPU replay used by tests/emul share one schedule.
 *
 * Exec must provide:
 *   void bounds();                         // bracket [lo0,hi0] per pencil
 *   void round(int r);                     // one multisection round
 *   void prepare(int buf);                 // hand brackets to the refinement
 *   void factor(int iter);                 // F pass
 *   void back(int corr_now, int corr_next);// B pass
 *   void check(int allow);                 // convergence marks
 *   void zero_counter(int which);          // which: 0 open brackets (also clears 2), 1 unconverged
 *   int  read_counter(int which);          // blocking read; 2 = open brackets that do not isolate yet
 */
#ifndef BSP_DRIVER_H
#define BSP_DRIVER_H

struct BspSchedule {
    int max_rounds;  /* multisection rounds cap                       */
    int min_iters;   /* refinement iterations always done (>= 3)      */
    int max_iters;   /* cap                                           */
    int first_check_round; /* first round after which the host polls  */
    int check_every;       /* ... and then every so many rounds (finished
                              brackets make a surplus round nearly free) */
    int open_ok;           /* hand over with this many brackets still open, provided each
                              isolates its eigenvalue: the refinement keeps bracketing
                              with the inertia of its own factorisations */
};

struct BspRunStats {
    int rounds;
    int iters;
    int brackets_open; /* brackets not narrowed when the cap was hit  */
    int unconverged;   /* eigenpairs above conv_tol at the end        */
};

template <class Exec>
inline BspRunStats bsp_run_chunk(Exec &ex, const BspSchedule &sch)
{
    BspRunStats st = {0, 0, 0, 0};
    ex.bounds();
    int r = 0;
    for (;;) {
        ex.zero_counter(0);
        ex.round(r);
        ++r;
        const int ce = sch.check_every > 0 ? sch.check_every : 1;
        if ((r >= sch.first_check_round && (r - sch.first_check_round) % ce == 0) || r >= sch.max_rounds) {
            st.brackets_open = ex.read_counter(0);
            if (st.brackets_open == 0 || r >= sch.max_rounds) break;
            if (st.brackets_open <= sch.open_ok && ex.read_counter(2) == 0) break;
        }
    }
    st.rounds = r;
    ex.prepare(r & 1);
    /* iteration t: plain for t < 2, residual-correction form afterwards */
    int t = 0;
    for (;;) {
        const int corr_now = (t >= 2), corr_next = (t + 1 >= 2);
        ex.factor(t);
        ex.back(corr_now, corr_next);
        ++t;
        if (t >= sch.min_iters || t >= sch.max_iters) {
            ex.zero_counter(1);
            ex.check(1);
            st.unconverged = ex.read_counter(1);
            if (st.unconverged == 0 || t >= sch.max_iters) break;
        }
    }
    st.iters = t;
    return st;
}

#endif
