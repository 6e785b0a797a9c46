/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * cyrk_oracle.c -- a plain-C, single-threaded CPU restatement ("port") of the
 * reference's adaptive explicit Runge-Kutta path, used ONLY as the parity
 * checker for the CUDA kernels (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline leg).  Nothing under cyrk_b200/ links, imports or calls it.
 *
 * PARITY PIN: this restatement is pinned against (i) the unmodified reference
 * compiled from /root/reference into oracle/_ref/libcyrk_ref.so (bit-for-bit
 * on t, y, events, counters -- tests/test_oracle_vs_ref.py, and the committed
 * fixtures under tests/golden/ generated from it by tests/golden/make_golden.py)
 * and (ii) the reference's own known-answer tests (dense KAT, brentq KAT,
 * closed-form accuracy test, status-code pins) -- tests/test_oracle_kat.py.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/CyRK).  Arithmetic is binary64, separately rounded (compile
 * with -ffp-contract=off, no -mfma), summed left to right in the reference's
 * term order with the same zero tableau entries skipped.
 *
 * Structure differs from the reference on purpose (table-driven tableaux, one
 * C struct, no classes); behaviour -- including the quirks listed in
 * SURVEY.md 8a -- does not.
 */
#include "oracle_api.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- status codes, cy/c_common.hpp:25-82 -------------------------------- */
enum {
    ST_CONVERGED = 20, ST_INITIALIZING = 10, ST_EVENT_TERMINATED = 2, ST_SUCCESS = 1, ST_NO_ERROR = 0,
    ST_ATTRIBUTE_ERROR = -7, ST_ARGUMENT_ERROR = -10, ST_DENSE_NOT_SAVED = -21, ST_BAD_CONFIG = -22,
    ST_OPT_SIGN = -30, ST_OPT_CONVERGENCE = -31, ST_MEMORY = -40,
    ST_MAX_STEPS_USER = -52, ST_MAX_STEPS_SYSARCH = -53, ST_STEP_SPACING = -54, ST_STEP_ACCEPT = -55,
    ST_NOT_SUCCESSFUL = -57, ST_BAD_FIRST_STEP = -80
};

/* ---- step-controller constants, cy/c_common.hpp:189-196 ------------------ */
static const double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0;
#define ORC_EPS 2.220446049250313e-16
static const double EPS_100 = ORC_EPS * 100.0;
/* cy/cysolver.hpp:44-46 */
static const double BRENTQ_XTOL = 4.0 * ORC_EPS, BRENTQ_RTOL = 4.0 * ORC_EPS;
#define BRENTQ_MAXITER 100

/* =========================================================================
 * Tableaux.  A row is a list of (stage index, coefficient) pairs in the order
 * the reference sums them; absent pairs are the zero entries it skips.
 * ========================================================================= */
typedef struct { int j; double c; } term_t;
typedef struct { const term_t* t; int n; } row_t;
#define ROW(a) { a, (int)(sizeof(a) / sizeof(a[0])) }

/* --- RK23 (Bogacki-Shampine), cy/rk.cpp:695-862, cy/rk.hpp:143-150 -------- */
static const double rk23_C[3] = { 0.0, 1.0 / 2.0, 3.0 / 4.0 };
static const term_t rk23_a1[] = { {0, 1.0 / 2.0} };
static const term_t rk23_a2[] = { {1, 3.0 / 4.0} };
static const row_t  rk23_A[3] = { {0, 0}, ROW(rk23_a1), ROW(rk23_a2) };
static const term_t rk23_b[] = { {0, 2.0 / 9.0}, {1, 1.0 / 3.0}, {2, 4.0 / 9.0} };
static const term_t rk23_e[] = { {0, 5.0 / 72.0}, {1, -1.0 / 12.0}, {2, -1.0 / 9.0}, {3, 1.0 / 8.0} };
static const term_t rk23_p1[] = { {0, -4.0 / 3.0}, {1, 1.0}, {2, 4.0 / 3.0}, {3, -1.0} };
static const term_t rk23_p2[] = { {0, 5.0 / 9.0}, {1, -2.0 / 3.0}, {2, -8.0 / 9.0}, {3, 1.0} };
static const row_t  rk23_P[3] = { {0, 0}, ROW(rk23_p1), ROW(rk23_p2) };

/* --- RK45 (Dormand-Prince 5(4)), cy/rk.cpp:886-1131, cy/rk.hpp:168-175 ---- */
static const double rk45_C[6] = { 0.0, 1.0 / 5.0, 3.0 / 10.0, 4.0 / 5.0, 8.0 / 9.0, 1.0 };
static const term_t rk45_a1[] = { {0, 1.0 / 5.0} };
static const term_t rk45_a2[] = { {0, 3.0 / 40.0}, {1, 9.0 / 40.0} };
static const term_t rk45_a3[] = { {0, 44.0 / 45.0}, {1, -56.0 / 15.0}, {2, 32.0 / 9.0} };
static const term_t rk45_a4[] = { {0, 19372.0 / 6561.0}, {1, -25360.0 / 2187.0}, {2, 64448.0 / 6561.0}, {3, -212.0 / 729.0} };
static const term_t rk45_a5[] = { {0, 9017.0 / 3168.0}, {1, -355.0 / 33.0}, {2, 46732.0 / 5247.0}, {3, 49.0 / 176.0}, {4, -5103.0 / 18656.0} };
static const row_t  rk45_A[6] = { {0, 0}, ROW(rk45_a1), ROW(rk45_a2), ROW(rk45_a3), ROW(rk45_a4), ROW(rk45_a5) };
static const term_t rk45_b[] = { {0, 35.0 / 384.0}, {2, 500.0 / 1113.0}, {3, 125.0 / 192.0}, {4, -2187.0 / 6784.0}, {5, 11.0 / 84.0} };
static const term_t rk45_e[] = { {0, -71.0 / 57600.0}, {2, 71.0 / 16695.0}, {3, -71.0 / 1920.0}, {4, 17253.0 / 339200.0}, {5, -22.0 / 525.0}, {6, 1.0 / 40.0} };
static const term_t rk45_p1[] = { {0, -8048581381.0 / 2820520608.0}, {2, 131558114200.0 / 32700410799.0}, {3, -1754552775.0 / 470086768.0},
                                  {4, 127303824393.0 / 49829197408.0}, {5, -282668133.0 / 205662961.0}, {6, 40617522.0 / 29380423.0} };
static const term_t rk45_p2[] = { {0, 8663915743.0 / 2820520608.0}, {2, -68118460800.0 / 10900136933.0}, {3, 14199869525.0 / 1410260304.0},
                                  {4, -318862633887.0 / 49829197408.0}, {5, 2019193451.0 / 616988883.0}, {6, -110615467.0 / 29380423.0} };
static const term_t rk45_p3[] = { {0, -12715105075.0 / 11282082432.0}, {2, 87487479700.0 / 32700410799.0}, {3, -10690763975.0 / 1880347072.0},
                                  {4, 701980252875.0 / 199316789632.0}, {5, -1453857185.0 / 822651844.0}, {6, 69997945.0 / 29380423.0} };
static const row_t  rk45_P[4] = { {0, 0}, ROW(rk45_p1), ROW(rk45_p2), ROW(rk45_p3) };

/* --- DOP853, cy/rk.cpp:1154-1676, cy/rk.hpp:194-205 ----------------------- */
static const double dop_C[12] = {
    0.0,
    0.526001519587677318785587544488e-01, 0.789002279381515978178381316732e-01,
    0.118350341907227396726757197510,     0.281649658092772603273242802490,
    0.333333333333333333333333333333,     0.25,
    0.307692307692307692307692307692,     0.651282051282051282051282051282,
    0.6,                                   0.857142857142857142857142857142,
    1.0 };
static const term_t dop_a1[] = { {0, 5.26001519587677318785587544488e-2} };
static const term_t dop_a2[] = { {0, 1.97250569845378994544595329183e-2}, {1, 5.91751709536136983633785987549e-2} };
static const term_t dop_a3[] = { {0, 2.95875854768068491816892993775e-2}, {2, 8.87627564304205475450678981324e-2} };
static const term_t dop_a4[] = { {0, 2.41365134159266685502369798665e-1}, {2, -8.84549479328286085344864962717e-1}, {3, 9.24834003261792003115737966543e-1} };
static const term_t dop_a5[] = { {0, 3.7037037037037037037037037037e-2}, {3, 1.70828608729473871279604482173e-1}, {4, 1.25467687566822425016691814123e-1} };
static const term_t dop_a6[] = { {0, 3.7109375e-2}, {3, 1.70252211019544039314978060272e-1}, {4, 6.02165389804559606850219397283e-2}, {5, -1.7578125e-2} };
static const term_t dop_a7[] = { {0, 3.70920001185047927108779319836e-2}, {3, 1.70383925712239993810214054705e-1}, {4, 1.07262030446373284651809199168e-1},
                                 {5, -1.53194377486244017527936158236e-2}, {6, 8.27378916381402288758473766002e-3} };
static const term_t dop_a8[] = { {0, 6.24110958716075717114429577812e-1}, {3, -3.36089262944694129406857109825}, {4, -8.68219346841726006818189891453e-1},
                                 {5, 2.75920996994467083049415600797e1}, {6, 2.01540675504778934086186788979e1}, {7, -4.34898841810699588477366255144e1} };
static const term_t dop_a9[] = { {0, 4.77662536438264365890433908527e-1}, {3, -2.48811461997166764192642586468}, {4, -5.90290826836842996371446475743e-1},
                                 {5, 2.12300514481811942347288949897e1}, {6, 1.52792336328824235832596922938e1}, {7, -3.32882109689848629194453265587e1},
                                 {8, -2.03312017085086261358222928593e-2} };
static const term_t dop_a10[] = { {0, -9.3714243008598732571704021658e-1}, {3, 5.18637242884406370830023853209}, {4, 1.09143734899672957818500254654},
                                  {5, -8.14978701074692612513997267357}, {6, -1.85200656599969598641566180701e1}, {7, 2.27394870993505042818970056734e1},
                                  {8, 2.49360555267965238987089396762}, {9, -3.0467644718982195003823669022} };
static const term_t dop_a11[] = { {0, 2.27331014751653820792359768449}, {3, -1.05344954667372501984066689879e1}, {4, -2.00087205822486249909675718444},
                                  {5, -1.79589318631187989172765950534e1}, {6, 2.79488845294199600508499808837e1}, {7, -2.85899827713502369474065508674},
                                  {8, -8.87285693353062954433549289258}, {9, 1.23605671757943030647266201528e1}, {10, 6.43392746015763530355970484046e-1} };
static const row_t dop_A[12] = { {0, 0}, ROW(dop_a1), ROW(dop_a2), ROW(dop_a3), ROW(dop_a4), ROW(dop_a5), ROW(dop_a6),
                                 ROW(dop_a7), ROW(dop_a8), ROW(dop_a9), ROW(dop_a10), ROW(dop_a11) };
static const term_t dop_b[] = { {0, 5.42937341165687622380535766363e-2}, {5, 4.45031289275240888144113950566}, {6, 1.89151789931450038304281599044},
                                {7, -5.8012039600105847814672114227}, {8, 3.1116436695781989440891606237e-1}, {9, -1.52160949662516078556178806805e-1},
                                {10, 2.01365400804030348374776537501e-1}, {11, 4.47106157277725905176885569043e-2} };
/* E3 = B - (bhh terms) folded at compile time exactly like cy/rk.cpp:1450-1457 */
static const term_t dop_e3[] = { {0, (5.42937341165687622380535766363e-2 - 0.244094488188976377952755905512)}, {5, 4.45031289275240888144113950566},
                                 {6, 1.89151789931450038304281599044}, {7, -5.8012039600105847814672114227},
                                 {8, (3.1116436695781989440891606237e-1 - 0.733846688281611857341361741547)}, {9, -1.52160949662516078556178806805e-1},
                                 {10, 2.01365400804030348374776537501e-1}, {11, (4.47106157277725905176885569043e-2 - 0.220588235294117647058823529412e-1)} };
static const term_t dop_e5[] = { {0, 0.1312004499419488073250102996e-1}, {5, -0.1225156446376204440720569753e+1}, {6, -0.4957589496572501915214079952},
                                 {7, 0.1664377182454986536961530415e+1}, {8, -0.3503288487499736816886487290}, {9, 0.3341791187130174790297318841},
                                 {10, 0.8192320648511571246570742613e-1}, {11, -0.2235530786388629525884427845e-1} };
/* extra stages for the dense interpolant, cy/rk.cpp:1506-1583 */
static const double dop_CX[3] = { 0.1, 0.2, 0.777777777777777777777777777778 };
static const term_t dop_x13[] = { {0, 5.61675022830479523392909219681e-2}, {6, 2.53500210216624811088794765333e-1}, {7, -2.46239037470802489917441475441e-1},
                                  {8, -1.24191423263816360469010140626e-1}, {9, 1.5329179827876569731206322685e-1}, {10, 8.20105229563468988491666602057e-3},
                                  {11, 7.56789766054569976138603589584e-3}, {12, -8.298e-3} };
static const term_t dop_x14[] = { {0, 3.18346481635021405060768473261e-2}, {5, 2.83009096723667755288322961402e-2}, {6, 5.35419883074385676223797384372e-2},
                                  {7, -5.49237485713909884646569340306e-2}, {10, -1.08347328697249322858509316994e-4}, {11, 3.82571090835658412954920192323e-4},
                                  {12, -3.40465008687404560802977114492e-4} };
static const double dop_x14_k13 = 1.41312443674632500278074618366e-1;
static const term_t dop_x15[] = { {0, -4.28896301583791923408573538692e-1}, {5, -4.69762141536116384314449447206}, {6, 7.68342119606259904184240953878},
                                  {7, 4.06898981839711007970213554331}, {8, 3.56727187455281109270669543021e-1}, {12, -1.39902416515901462129418009734e-3} };
static const double dop_x15_k13 = 2.9475147891527723389556272149;
static const double dop_x15_k14 = -9.15095847217987001081870187138;
static const term_t dop_d0[] = { {0, -0.84289382761090128651353491142e+1}, {5, 0.56671495351937776962531783590}, {6, -0.30689499459498916912797304727e+1},
                                 {7, 0.23846676565120698287728149680e+1}, {8, 0.21170345824450282767155149946e+1}, {9, -0.87139158377797299206789907490},
                                 {10, 0.22404374302607882758541771650e+1}, {11, 0.63157877876946881815570249290}, {12, -0.88990336451333310820698117400e-1},
                                 {13, 0.18148505520854727256656404962e+2}, {14, -0.91946323924783554000451984436e+1}, {15, -0.44360363875948939664310572000e+1} };
static const term_t dop_d1[] = { {0, 0.10427508642579134603413151009e+2}, {5, 0.24228349177525818288430175319e+3}, {6, 0.16520045171727028198505394887e+3},
                                 {7, -0.37454675472269020279518312152e+3}, {8, -0.22113666853125306036270938578e+2}, {9, 0.77334326684722638389603898808e+1},
                                 {10, -0.30674084731089398182061213626e+2}, {11, -0.93321305264302278729567221706e+1}, {12, 0.15697238121770843886131091075e+2},
                                 {13, -0.31139403219565177677282850411e+2}, {14, -0.93529243588444783865713862664e+1}, {15, 0.35816841486394083752465898540e+2} };
static const term_t dop_d2[] = { {0, 0.19985053242002433820987653617e+2}, {5, -0.38703730874935176555105901742e+3}, {6, -0.18917813819516756882830838328e+3},
                                 {7, 0.52780815920542364900561016686e+3}, {8, -0.11573902539959630126141871134e+2}, {9, 0.68812326946963000169666922661e+1},
                                 {10, -0.10006050966910838403183860980e+1}, {11, 0.77771377980534432092869265740}, {12, -0.27782057523535084065932004339e+1},
                                 {13, -0.60196695231264120758267380846e+2}, {14, 0.84320405506677161018159903784e+2}, {15, 0.11992291136182789328035130030e+2} };
static const term_t dop_d3[] = { {0, -0.25693933462703749003312586129e+2}, {5, -0.15418974869023643374053993627e+3}, {6, -0.23152937917604549567536039109e+3},
                                 {7, 0.35763911791061412378285349910e+3}, {8, 0.93405324183624310003907691704e+2}, {9, -0.37458323136451633156875139351e+2},
                                 {10, 0.10409964950896230045147246184e+3}, {11, 0.29840293426660503123344363579e+2}, {12, -0.43533456590011143754432175058e+2},
                                 {13, 0.96324553959188282948394950600e+2}, {14, -0.39177261675615439165231486172e+2}, {15, -0.14972683625798562581422125276e+3} };
static const row_t dop_D[4] = { ROW(dop_d0), ROW(dop_d1), ROW(dop_d2), ROW(dop_d3) };

typedef struct {
    int method, n_stages, K_stride, Q_order;
    double err_exponent;
    const double* C;
    const row_t* A;
    row_t B, E;
    const row_t* P;
} tableau_t;

static const tableau_t TAB23 = { 3, 3, 4, 3, 1.0 / (2.0 + 1.0), rk23_C, rk23_A, ROW(rk23_b), ROW(rk23_e), rk23_P };
static const tableau_t TAB45 = { 4, 6, 7, 4, 1.0 / (4.0 + 1.0), rk45_C, rk45_A, ROW(rk45_b), ROW(rk45_e), rk45_P };
/* K_stride 18 = 13 K + 3 extra + 2 accumulators, cy/rk.cpp:1148 */
static const tableau_t TAB853 = { 5, 12, 18, 7, 1.0 / (7.0 + 1.0), dop_C, dop_A, ROW(dop_b), {0, 0}, 0 };

static inline double row_dot(const row_t r, const double* k)
{
    double acc = r.t[0].c * k[r.t[0].j];
    for (int i = 1; i < r.n; i++) acc += r.t[i].c * k[r.t[i].j];
    return acc;
}

/* =========================================================================
 * Solver state (CySolverBase + RKSolver + CySolverResult + CySolverDense rolled
 * into one struct).
 * ========================================================================= */
typedef struct {
    double t_old, t_now, step;
    double* state;                 /* [y_old (ny) | Q (ny * Q_order)], cy/dense.cpp:27-37 */
} dense_t;

typedef struct {
    orc_event_fn fn;
    size_t max_allowed, count;
    int direction, is_active;
    double last_root, g_old;
    double* y_at_root;             /* ndy */
    double* t_hits; double* y_hits; size_t n_hits, cap;
} event_t;

typedef struct {
    const tableau_t* tab;
    orc_diffeq_fn diffeq; orc_pre_eval_fn pre_eval;
    size_t ny, nextra, ndy;
    double ny_dbl, ny_sqrt;
    char* args; size_t args_len;
    double* rtols; double* atols; size_t n_rtols, n_atols; int arr_rtol, arr_atol;
    /* state: holders laid out contiguously like cy/cysolver.cpp:276-301 */
    double* yh; double* dyh;
    double *y_old, *y_now, *y_tmp, *y_interp, *dy_old, *dy_now, *dy_tmp, *dy_tmp2;
    double t_start, t_end, t_delta_abs, dir_inf, t_old, t_now, t_tmp;
    int fwd;
    double* K;
    double step, step_size, max_step_size, first_step;
    size_t len_t, max_num_steps; int user_max_steps;
    int error_flag, setup_ok;
    /* t_eval */
    double* t_eval; size_t n_t_eval, t_eval_index_old; int use_t_eval, t_eval_finished;
    /* events */
    event_t* ev; size_t n_events; size_t* active; size_t n_active; double termination_root;
    double* root_y;                /* 2 * ndy, OptimizeInfo::y_vec */
    int event_terminated; size_t event_terminate_index;
    /* results (CySolverResult) */
    int status, success, dense_output, capture_extra, fix_quirk11;
    size_t size, steps_taken, num_interpolates, cap;
    double* t_out; double* y_out;
    dense_t* dense; size_t n_dense, dense_cap;
    double* interp_time; size_t n_interp_time, interp_cap;
    size_t nfev;
} solver_t;

static inline void call_diffeq(solver_t* s)   /* cy/cysolver.cpp:222-231 */
{
    s->nfev++;
    s->diffeq(s->dy_now, s->t_now, s->y_now, s->args, s->pre_eval);
}

static inline void offload_to_temp(solver_t* s)   /* cy/cysolver.cpp:260-266 */
{
    memcpy(s->y_tmp, s->y_now, sizeof(double) * s->ny);
    memcpy(s->dy_tmp, s->dy_now, sizeof(double) * s->ndy);
    s->t_tmp = s->t_now;
}
static inline void load_back_from_temp(solver_t* s)   /* cy/cysolver.cpp:268-274 */
{
    memcpy(s->y_now, s->y_tmp, sizeof(double) * s->ny);
    memcpy(s->dy_now, s->dy_tmp, sizeof(double) * s->ndy);
    s->t_now = s->t_tmp;
}

/* ---- storage (cy/cysolution.cpp:396-421); growth policy is irrelevant to results */
static void save_data(solver_t* s, double t, const double* y, const double* dy)
{
    if (s->size + 1 > s->cap) {
        s->cap = s->cap ? s->cap * 2 : 256;
        s->t_out = (double*)realloc(s->t_out, sizeof(double) * s->cap);
        s->y_out = (double*)realloc(s->y_out, sizeof(double) * s->cap * s->ndy);
    }
    s->t_out[s->size] = t;
    double* row = s->y_out + s->size * s->ndy;
    memcpy(row, y, sizeof(double) * s->ny);
    if (s->capture_extra) memcpy(row + s->ny, dy + s->ny, sizeof(double) * s->nextra);
    s->size++;
}

/* ---- stages: cy/rk.cpp:695-776 (RK23), 886-1025 (RK45), 1154-1412 (DOP853) */
static void compute_stages(solver_t* s)
{
    const tableau_t* T = s->tab;
    const size_t ny = s->ny; const int ns = T->n_stages; const int ks = T->K_stride;
    const double original_time = s->t_now;
    const double step = s->step;

    /* stage 1: k0 = dy_old (FSAL).  The three methods group the product differently:
       RK23 (step*k0)*a (rk.cpp:721); RK45/DOP853 step*(a*k0) (rk.cpp:908,1174). */
    s->t_now = s->t_old + T->C[1] * step;
    for (size_t i = 0; i < ny; i++) {
        double* k = s->K + i * ks;
        k[0] = s->dy_old[i];
        const double a = T->A[1].t[0].c;
        if (T->method == 3) s->y_now[i] = s->y_old[i] + (step * k[0] * a);
        else                s->y_now[i] = s->y_old[i] + step * (a * k[0]);
    }
    call_diffeq(s);

    for (int st = 2; st < ns; st++) {
        s->t_now = s->t_old + T->C[st] * step;
        for (size_t i = 0; i < ny; i++) {
            double* k = s->K + i * ks;
            k[st - 1] = s->dy_now[i];
            s->y_now[i] = s->y_old[i] + (step * row_dot(T->A[st], k));
        }
        call_diffeq(s);
    }

    /* y_new = y_old + step * (B . K); then the FSAL evaluation at (t_new, y_new) */
    s->t_now = original_time;
    for (size_t i = 0; i < ny; i++) {
        double* k = s->K + i * ks;
        k[ns - 1] = s->dy_now[i];
        s->y_now[i] = s->y_old[i] + (step * row_dot(T->B, k));
    }
    call_diffeq(s);
    for (size_t i = 0; i < ny; i++) s->K[i * ks + ns] = s->dy_now[i];
}

/* ---- error norm: cy/rk.cpp:778-823 (RK23), 1027-1075 (RK45), 1414-1489 (DOP853) */
static double estimate_error(solver_t* s)
{
    const tableau_t* T = s->tab;
    const int ks = T->K_stride;
    double rtol = s->rtols[0], atol = s->atols[0];
    if (T->method != 5) {
        double norm = 0.0;
        for (size_t i = 0; i < s->ny; i++) {
            rtol = s->arr_rtol ? s->rtols[i] : rtol;
            atol = s->arr_atol ? s->atols[i] : atol;
            const double* k = s->K + i * ks;
            const double e = row_dot(T->E, k);
            const double sc = e / (atol + fmax(fabs(s->y_old[i]), fabs(s->y_now[i])) * rtol);
            norm += (sc * sc);
        }
        return s->step_size * sqrt(norm) / s->ny_sqrt;
    }
    double n3 = 0.0, n5 = 0.0;
    const row_t E3 = ROW(dop_e3), E5 = ROW(dop_e5);
    for (size_t i = 0; i < s->ny; i++) {
        rtol = s->arr_rtol ? s->rtols[i] : rtol;
        atol = s->arr_atol ? s->atols[i] : atol;
        const double* k = s->K + i * ks;
        const double scale_inv = 1.0 / (atol + fmax(fabs(s->y_old[i]), fabs(s->y_now[i])) * rtol);
        const double e3 = scale_inv * (row_dot(E3, k));
        const double e5 = scale_inv * (row_dot(E5, k));
        n3 += (e3 * e3);
        n5 += (e5 * e5);
    }
    if ((n5 == 0.0) && (n3 == 0.0)) return 0.0;
    const double denom = n5 + 0.01 * n3;
    return s->step_size * n5 / sqrt(denom * s->ny_dbl);
}

/* ---- first step heuristic: cy/rk.cpp:183-282 (no clamp to the span, quirk 2) */
static void calc_first_step_size(solver_t* s)
{
    double rtol = s->rtols[0], atol = s->atols[0];
    double d0 = 0.0, d1 = 0.0;
    for (size_t i = 0; i < s->ny; i++) {
        rtol = s->arr_rtol ? s->rtols[i] : rtol;
        atol = s->arr_atol ? s->atols[i] : atol;
        const double yo = s->y_old[i];
        const double scale = atol + fabs(yo) * rtol;
        const double a0 = yo / scale;
        const double a1 = s->dy_old[i] / scale;
        d0 += (a0 * a0);
        d1 += (a1 * a1);
    }
    d0 = sqrt(d0) / s->ny_sqrt;
    d1 = sqrt(d1) / s->ny_sqrt;
    double h0 = 1.0e-6;
    if (!((d0 < 1.0e-5) || (d1 < 1.0e-5))) h0 = 0.01 * d0 / d1;
    const double h0_dir = s->fwd ? h0 : -h0;

    s->t_now = s->t_old + h0_dir;
    for (size_t i = 0; i < s->ny; i++) s->y_now[i] = s->y_old[i] + h0_dir * s->dy_old[i];
    call_diffeq(s);

    double d2 = 0.0;
    for (size_t i = 0; i < s->ny; i++) {
        if (s->arr_rtol) rtol = s->rtols[i];
        if (s->arr_atol) atol = s->atols[i];
        const double scale = atol + fabs(s->y_old[i]) * rtol;
        const double a2 = (s->dy_now[i] - s->dy_old[i]) / scale;
        d2 += (a2 * a2);
    }
    d2 = sqrt(d2) / (h0 * s->ny_sqrt);
    double h1;
    if ((d1 <= 1.0e-15) && (d2 <= 1.0e-15)) h1 = fmax(1.0e-6, h0 * 1.0e-3);
    else h1 = pow((0.01 / fmax(d1, d2)), s->tab->err_exponent);
    s->step_size = fmax(10. * fabs(nextafter(s->t_old, s->dir_inf) - s->t_old), fmin(100.0 * h0, h1));
}

/* ---- one accepted step = loop of attempts: cy/rk.cpp:426-538 */
static void step_implementation(solver_t* s)
{
    const double ee = s->tab->err_exponent;
    const double min_step = 10. * fabs(nextafter(s->t_old, s->dir_inf) - s->t_old);
    /* std::clamp(v, lo, hi): v<lo ? lo : (hi<v ? hi : v) */
    if (s->step_size < min_step) s->step_size = min_step;
    else if (s->max_step_size < s->step_size) s->step_size = s->max_step_size;

    int accepted = 0, rejected = 0;
    while (!accepted) {
        if (s->step_size < min_step) { s->status = ST_STEP_SPACING; return; }
        double over;
        if (s->fwd) { s->step = s->step_size;  s->t_now = s->t_old + s->step; over = s->t_now - s->t_end; }
        else        { s->step = -s->step_size; s->t_now = s->t_old + s->step; over = s->t_end - s->t_now; }
        if (over > 0.0) {
            s->t_now = s->t_end;
            s->step = s->t_now - s->t_old;
            s->step_size = s->fwd ? s->step : -s->step;
        }
        compute_stages(s);
        const double err = estimate_error(s);
        const double err_safe = SAFETY / pow(err, ee);     /* a division (quirk 1) */
        if (err < 1.0) {
            double f = MAX_FACTOR;
            if (err != 0.0) f = fmin(f, err_safe);
            if (rejected)   f = fmin(f, 1.0);
            s->step_size *= f;
            accepted = 1;
        } else {
            s->step_size *= fmax(MIN_FACTOR, err_safe);
            rejected = 1;
        }
    }
}

/* ---- dense interpolant construction: cy/dense.cpp:47-69 + set_Q_array
 *      (cy/rk.cpp:825-862 RK23, 1077-1131 RK45, 1491-1676 DOP853) */
static void set_Q(solver_t* s, double* Q)
{
    const tableau_t* T = s->tab;
    const int ks = T->K_stride; const int qo = T->Q_order;
    if (T->method != 5) {
        for (size_t i = 0; i < s->ny; i++) {
            const double* k = s->K + i * ks;
            Q[i * qo] = k[0];
            for (int c = 1; c < qo; c++) Q[i * qo + c] = row_dot(T->P[c], k);
        }
        return;
    }
    /* DOP853: three more stages with running accumulators in k[16], k[17] */
    offload_to_temp(s);
    const double step = s->step;
    const row_t X13 = ROW(dop_x13), X14 = ROW(dop_x14), X15 = ROW(dop_x15);
    for (size_t i = 0; i < s->ny; i++) {
        double* k = s->K + i * ks;
        const double acc13 = row_dot(X13, k);
        k[16] = row_dot(X14, k);
        k[17] = row_dot(X15, k);
        s->y_now[i] = s->y_old[i] + acc13 * step;
    }
    s->t_now = s->t_old + (step * dop_CX[0]);
    call_diffeq(s);
    for (size_t i = 0; i < s->ny; i++) {
        double* k = s->K + i * ks;
        k[13] = s->dy_now[i];
        k[16] += k[13] * dop_x14_k13;
        k[17] += k[13] * dop_x15_k13;
        s->y_now[i] = s->y_old[i] + k[16] * step;
    }
    s->t_now = s->t_old + (step * dop_CX[1]);
    call_diffeq(s);
    for (size_t i = 0; i < s->ny; i++) {
        double* k = s->K + i * ks;
        k[14] = s->dy_now[i];
        k[17] += k[14] * dop_x15_k14;
        s->y_now[i] = s->y_old[i] + k[17] * step;
    }
    s->t_now = s->t_old + (step * dop_CX[2]);
    call_diffeq(s);
    for (size_t i = 0; i < s->ny; i++) {
        double* k = s->K + i * ks;
        k[15] = s->dy_now[i];
        double* q = Q + i * 7;
        /* F rows stored reversed (rk.cpp:1660-1671) */
        q[0] = step * row_dot(dop_D[3], k);
        q[1] = step * row_dot(dop_D[2], k);
        q[2] = step * row_dot(dop_D[1], k);
        q[3] = step * row_dot(dop_D[0], k);
        const double delta_y = s->y_tmp[i] - s->y_old[i];
        const double sum_dy  = s->dy_tmp[i] + k[0];
        q[4] = 2.0 * delta_y - step * sum_dy;
        q[5] = step * k[0] - delta_y;
        q[6] = delta_y;
    }
    load_back_from_temp(s);
}

static void dense_set_state(solver_t* s, dense_t* d)
{
    d->t_old = s->t_old;
    d->t_now = s->t_now;
    d->step = d->t_now - d->t_old;
    memcpy(d->state, s->y_old, sizeof(double) * s->ny);
    set_Q(s, d->state + s->ny);
}

/* ---- interpolant evaluation: cy/dense.cpp:71-207 */
static void dense_call(solver_t* s, const dense_t* d, double t, double* out)
{
    const double x = (t - d->t_old) / d->step;
    const double* ys = d->state;
    const double* Q = d->state + s->ny;
    const int m = s->tab->method;
    for (size_t i = 0; i < s->ny; i++) {
        if (m == 3 || m == 4) {
            const int qo = s->tab->Q_order;
            const double* q = Q + qo * i;
            double cp = x;
            double acc = q[0] * cp;
            for (int p = 1; p < qo; p++) { cp *= x; acc += q[p] * cp; }
            acc *= d->step;
            out[i] = ys[i] + acc;
        } else {
            const double* q = Q + 7 * i;
            double acc = q[0];
            acc *= x;
            acc += q[1]; acc *= (1.0 - x);
            acc += q[2]; acc *= x;
            acc += q[3]; acc *= (1.0 - x);
            acc += q[4]; acc *= x;
            acc += q[5]; acc *= (1.0 - x);
            acc += q[6]; acc *= x;
            out[i] = ys[i] + acc;
        }
    }
    if (s->capture_extra) {
        /* extras are not interpolated: re-evaluate the RHS at (t, y_interp), dense.cpp:172-206 */
        offload_to_temp(s);
        memmove(s->y_now, out, sizeof(double) * s->ny);    /* out may BE y_now (termination path) */
        s->t_now = t;
        call_diffeq(s);
        for (size_t i = s->ny; i < s->ndy; i++) out[i] = s->dy_now[i];
        load_back_from_temp(s);
    }
}

/* ---- CySolverResult::build_dense, cy/cysolution.cpp:450-482 */
static void build_dense(solver_t* s, int save)
{
    if (save) {
        s->num_interpolates++;
        if (s->n_dense + 1 > s->dense_cap) {
            s->dense_cap = s->dense_cap ? s->dense_cap * 2 : 256;
            s->dense = (dense_t*)realloc(s->dense, sizeof(dense_t) * s->dense_cap);
        }
        dense_t* d = &s->dense[s->n_dense++];
        d->state = (double*)malloc(sizeof(double) * s->ny * (s->tab->Q_order + 1));
        dense_set_state(s, d);
        if (s->use_t_eval) {
            if (s->n_interp_time + 1 > s->interp_cap) {
                s->interp_cap = s->interp_cap ? s->interp_cap * 2 : 256;
                s->interp_time = (double*)realloc(s->interp_time, sizeof(double) * s->interp_cap);
            }
            s->interp_time[s->n_interp_time++] = s->t_now;
        }
    } else {
        dense_set_state(s, &s->dense[0]);
    }
}

/* ---- Brent root finder: optimize/c_brentq.cpp:41-204.  `eval(ctx, x, buf)`
 *      evaluates the interpolant into y-buffer `buf` (0 or 1) and returns g. */
typedef double (*brent_eval_fn)(void* ctx, double x, int buf);
typedef struct { int error_num; size_t funcalls, iterations; int y_at_root_buf; } brent_info;

static double brentq(brent_eval_fn eval, void* ctx, double xa, double xb, double xtol, double rtol,
                     size_t iter, brent_info* info)
{
    double xpre = xa, xcur = xb;
    double xblk = 0., fblk = 0., spre = 0., scur = 0.;
    double stry, dpre, dblk;
    info->error_num = ST_INITIALIZING;
    info->y_at_root_buf = 0;
    double fpre = eval(ctx, xpre, 0);
    double fcur = eval(ctx, xcur, 1);
    info->funcalls = 2;
    if (fpre == 0) { info->error_num = ST_CONVERGED; return xpre; }
    if (fcur == 0) { info->error_num = ST_CONVERGED; info->y_at_root_buf = 1; return xcur; }
    if ((signbit(fpre) != 0) == (signbit(fcur) != 0)) { info->error_num = ST_OPT_SIGN; return 0.0; }
    info->iterations = 0;
    for (size_t i = 0; i < iter; i++) {
        info->iterations++;
        if ((fpre != 0) && (fcur != 0) && ((signbit(fpre) != 0) != (signbit(fcur) != 0))) {
            xblk = xpre; fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        const double delta = (xtol + rtol * fabs(xcur)) / 2.0;
        const double sbis = (xblk - xcur) / 2.0;
        if ((fcur == 0) || (fabs(sbis) < delta)) { info->error_num = ST_CONVERGED; return xcur; }
        if ((fabs(spre) > delta) && (fabs(fcur) < fabs(fpre))) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                dpre = (fpre - fcur) / (xpre - xcur);
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            const double m_a = fabs(spre), m_b = 3.0 * fabs(sbis) - delta;
            if (2.0 * fabs(stry) < (m_a < m_b ? m_a : m_b)) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        /* always lands in buffer 0 -- the returned y may lag the returned x by one iterate (quirk 5) */
        fcur = eval(ctx, xcur, 0);
        info->funcalls++;
    }
    info->error_num = ST_OPT_CONVERGENCE;
    return xcur;
}

typedef struct { solver_t* s; event_t* e; const dense_t* d; } ev_ctx;
static double ev_eval(void* c, double x, int buf)
{
    ev_ctx* k = (ev_ctx*)c;
    double* y = k->s->root_y + (size_t)buf * k->s->ndy;
    dense_call(k->s, k->d, x, y);
    return k->e->fn(x, y, k->s->args);
}

/* ---- event detection: cy/cysolver.cpp:563-750 */
static int check_events(solver_t* s)
{
    const dense_t* d = &s->dense[s->n_dense - 1];
    double* y_use = s->y_now;
    if (s->capture_extra) {
        memcpy(s->dy_tmp2, s->y_now, sizeof(double) * s->ny);
        memcpy(s->dy_tmp2 + s->ny, s->dy_now + s->ny, sizeof(double) * s->nextra);
        y_use = s->dy_tmp2;
    }
    s->n_active = 0;
    s->termination_root = s->dir_inf;
    for (size_t ei = 0; ei < s->n_events; ei++) {
        event_t* e = &s->ev[ei];
        const double g_now = e->fn(s->t_now, y_use, s->args);
        const double g_old = e->g_old;
        const int up = (g_old <= 0.0) && (g_now >= 0.0);
        const int down = (g_old >= 0.0) && (g_now <= 0.0);   /* 0 -> 0 triggers (quirk 4) */
        int trig = 0;
        if ((e->direction == 0) && (up || down)) trig = 1;
        else if ((e->direction > 0) && up) trig = 1;
        else if ((e->direction < 0) && down) trig = 1;
        e->g_old = g_now;
        if (!trig) { e->is_active = 0; continue; }
        e->is_active = 1;
        e->count++;
        s->active[s->n_active++] = ei;
        brent_info info = { ST_NO_ERROR, 0, 0, 0 };
        ev_ctx ctx = { s, e, d };
        e->last_root = brentq(ev_eval, &ctx, s->t_old, s->t_now, BRENTQ_XTOL, BRENTQ_RTOL, BRENTQ_MAXITER, &info);
        if (info.error_num != ST_CONVERGED) { s->error_flag = 1; return info.error_num; }
        memcpy(e->y_at_root, s->root_y + (size_t)info.y_at_root_buf * s->ndy, sizeof(double) * s->ndy);
        if (e->count >= e->max_allowed) {
            s->event_terminated = 1;
            if (s->fwd) { if (e->last_root < s->termination_root) s->termination_root = e->last_root; }
            else        { if (e->last_root > s->termination_root) s->termination_root = e->last_root; }
        }
    }
    if (s->event_terminated) {
        size_t w = 0;
        for (size_t r = 0; r < s->n_active; r++) {
            const size_t ei = s->active[r];
            event_t* e = &s->ev[ei];
            if (e->last_root == s->termination_root) s->event_terminate_index = ei;
            const int beyond = s->fwd ? (e->last_root > s->termination_root) : (e->last_root < s->termination_root);
            if (beyond) e->is_active = 0;
            else s->active[w++] = ei;
        }
        s->n_active = w;
    }
    return ST_NO_ERROR;
}

/* ---- cy/cysolution.cpp:423-448 */
static void record_event_data(solver_t* s)
{
    for (size_t r = 0; r < s->n_active; r++) {
        event_t* e = &s->ev[s->active[r]];
        if (e->n_hits + 1 > e->cap) {
            e->cap = e->cap ? e->cap * 2 : 64;
            e->t_hits = (double*)realloc(e->t_hits, sizeof(double) * e->cap);
            e->y_hits = (double*)realloc(e->y_hits, sizeof(double) * e->cap * s->ndy);
        }
        e->t_hits[e->n_hits] = e->last_root;
        memcpy(e->y_hits + e->n_hits * s->ndy, e->y_at_root, sizeof(double) * s->ndy);
        e->n_hits++;
    }
}

static size_t lower_bound_d(const double* a, size_t n, double v)
{ size_t lo = 0, hi = n; while (lo < hi) { size_t m = lo + (hi - lo) / 2; if (a[m] < v) lo = m + 1; else hi = m; } return lo; }
static size_t upper_bound_d(const double* a, size_t n, double v)
{ size_t lo = 0, hi = n; while (lo < hi) { size_t m = lo + (hi - lo) / 2; if (!(v < a[m])) lo = m + 1; else hi = m; } return lo; }

/* ---- one pass of the hot loop body: cy/cysolver.cpp:752-991 */
static void take_step(solver_t* s)
{
    if (s->t_now == s->t_end) {                  /* exact compare (quirk 3) */
        s->t_old = s->t_end;
        s->status = ST_SUCCESS;
    } else if (s->len_t >= s->max_num_steps) {
        s->error_flag = 1;
        s->status = s->user_max_steps ? ST_MAX_STEPS_USER : ST_MAX_STEPS_SYSARCH;
    } else {
        int do_save = 1, roll = 1, dense_built = 0;
        step_implementation(s);
        s->len_t++;
        s->steps_taken++;
        if (s->dense_output) { build_dense(s, 1); dense_built = 1; }
        if (s->n_events) {
            if (!dense_built) { build_dense(s, 0); dense_built = 1; }
            const int ev_status = check_events(s);
            s->status = ev_status;               /* update_status(p_check_events()) */
            record_event_data(s);
            if (s->event_terminated && !s->error_flag) {
                s->status = ST_EVENT_TERMINATED;
                if (s->t_now != s->termination_root) {
                    s->t_now = s->termination_root;
                    const dense_t* d = &s->dense[s->n_dense - 1];
                    if (s->fix_quirk11 && s->capture_extra) {
                        /* corrected behaviour: interpolate into a num_dy-sized scratch, keep only y */
                        dense_call(s, d, s->t_now, s->root_y);
                        memcpy(s->y_now, s->root_y, sizeof(double) * s->ny);
                    } else {
                        /* reference: writes num_dy values through y_now_ptr (spills into y_tmp, quirk 11) */
                        dense_call(s, d, s->t_now, s->y_now);
                    }
                    if (s->capture_extra) call_diffeq(s);
                }
            }
        }
        if (s->use_t_eval && !s->t_eval_finished && !s->error_flag) {
            do_save = 0;
            const size_t lo = lower_bound_d(s->t_eval, s->n_t_eval, s->t_now);
            const size_t up = upper_bound_d(s->t_eval, s->n_t_eval, s->t_now);
            size_t idx_new;
            if (lo == up) idx_new = lo;
            else if (s->fwd) { idx_new = up; if (idx_new == s->n_t_eval) s->t_eval_finished = 1; }
            else             { idx_new = lo; if (idx_new == 0) s->t_eval_finished = 1; }
            size_t delta; int any;
            if (s->fwd) { any = idx_new > s->t_eval_index_old; delta = idx_new - s->t_eval_index_old; }
            else        { any = s->t_eval_index_old > idx_new; delta = s->t_eval_index_old - idx_new; }
            if (any) {
                if (!dense_built) { build_dense(s, 0); dense_built = 1; }
                if (s->capture_extra && !s->error_flag) {
                    s->t_old = s->t_now;
                    memcpy(s->y_old, s->y_now, sizeof(double) * s->ny);
                    memcpy(s->dy_old, s->dy_now, sizeof(double) * s->ndy);
                    roll = 0;
                }
                if (!s->error_flag && dense_built) {
                    const dense_t* d = &s->dense[s->n_dense - 1];
                    for (size_t i = 0; i < delta; i++) {
                        const double ti = s->fwd ? s->t_eval[s->t_eval_index_old + i]
                                                 : s->t_eval[s->t_eval_index_old - i - 1];
                        dense_call(s, d, ti, s->y_interp);
                        if (s->capture_extra) {
                            s->t_now = ti;                      /* t_now is left at the last t_eval point */
                            memcpy(s->y_now, s->y_interp, sizeof(double) * s->ny);
                            call_diffeq(s);
                        }
                        save_data(s, ti, s->y_interp, s->dy_now);
                    }
                }
            }
            s->t_eval_index_old = idx_new;
        }
        if (do_save && !s->error_flag) save_data(s, s->t_now, s->y_now, s->dy_now);
        if (roll && !s->error_flag) {
            s->t_old = s->t_now;
            memcpy(s->y_old, s->y_now, sizeof(double) * s->ny);
            memcpy(s->dy_old, s->dy_now, sizeof(double) * s->ndy);
        }
    }
    if (s->status == ST_SUCCESS || s->status == ST_EVENT_TERMINATED) s->success = 1;
}

/* ---- cy/c_common.cpp:20-56 */
static void find_max_num_steps(size_t ny, size_t nextra, size_t user, size_t ram_MB, int* user_flag, size_t* out)
{
    const double ram = ram_MB * (1000.0 * 1000.0) / (sizeof(double) * (1.0 + ny + nextra));
    const size_t ram_steps = (size_t)floor(ram);
    *user_flag = 0; *out = ram_steps;
    if (user > 0) {
        if (user > ram_steps) *out = ram_steps;
        else { *user_flag = 1; *out = user; }
    }
    if (*out > (SIZE_MAX / 10)) *out = SIZE_MAX / 10;
}

static void solver_free(solver_t* s)
{
    if (!s) return;
    free(s->args); free(s->rtols); free(s->atols); free(s->yh); free(s->dyh); free(s->K);
    free(s->t_eval); free(s->active); free(s->root_y); free(s->t_out); free(s->y_out); free(s->interp_time);
    for (size_t i = 0; i < s->n_dense; i++) free(s->dense[i].state);
    free(s->dense);
    for (size_t i = 0; i < s->n_events; i++) { free(s->ev[i].y_at_root); free(s->ev[i].t_hits); free(s->ev[i].y_hits); }
    free(s->ev);
    free(s);
}

/* ---- setup: cy/cysolution.cpp:220-393 -> cy/rk.cpp:543-623 -> cy/cysolver.cpp:303-543 */
static solver_t* solver_setup(const orc_problem* p)
{
    solver_t* s = (solver_t*)calloc(1, sizeof(solver_t));
    s->status = ST_INITIALIZING;
    s->tab = p->method == 3 ? &TAB23 : p->method == 4 ? &TAB45 : p->method == 5 ? &TAB853 : 0;
    s->event_terminate_index = SIZE_MAX;
    if (!s->tab) { s->status = -3; return s; }
    if (p->num_y == 0) { s->status = ST_BAD_CONFIG; return s; }
    s->diffeq = p->diffeq; s->pre_eval = p->pre_eval;
    s->ny = p->num_y; s->nextra = p->num_extra; s->ndy = s->ny + s->nextra;
    s->ny_dbl = (double)s->ny; s->ny_sqrt = sqrt(s->ny_dbl);
    s->capture_extra = s->nextra > 0;
    s->dense_output = p->dense_output != 0;
    s->fix_quirk11 = p->fix_quirk11;
    s->args_len = p->args_len;
    s->args = (char*)malloc(p->args_len ? p->args_len : 1);
    if (p->args_len) memcpy(s->args, p->args, p->args_len);
    s->n_events = p->n_events;

    /* --- CySolverBase::setup, cysolver.cpp:355-372 */
    s->t_start = p->t_start; s->t_end = p->t_end;
    const double t_delta = s->t_end - s->t_start;
    s->t_delta_abs = fabs(t_delta);
    s->fwd = t_delta >= 0.0;
    s->dir_inf = s->fwd ? INFINITY : -INFINITY;
    s->termination_root = s->dir_inf;
    s->n_t_eval = p->n_t_eval;
    s->use_t_eval = s->n_t_eval > 0;
    s->yh = (double*)calloc(4 * s->ny + s->ndy, sizeof(double));   /* + slack for the y_interp overrun (quirk 11 note) */
    s->dyh = (double*)calloc(4 * s->ndy, sizeof(double));
    s->y_old = s->yh; s->y_now = s->yh + s->ny; s->y_tmp = s->yh + 2 * s->ny; s->y_interp = s->yh + 3 * s->ny;
    s->dy_old = s->dyh; s->dy_now = s->dyh + s->ndy; s->dy_tmp = s->dyh + 2 * s->ndy; s->dy_tmp2 = s->dyh + 3 * s->ndy;
    if (s->use_t_eval) {
        s->t_eval = (double*)malloc(sizeof(double) * s->n_t_eval);
        if (s->fwd) memcpy(s->t_eval, p->t_eval, sizeof(double) * s->n_t_eval);
        else for (size_t i = 0; i < s->n_t_eval; i++) s->t_eval[i] = p->t_eval[s->n_t_eval - 1 - i];  /* :381-400 */
    }
    find_max_num_steps(s->ny, s->nextra, p->max_num_steps, p->max_ram_MB, &s->user_max_steps, &s->max_num_steps);
    s->root_y = (double*)calloc(2 * s->ndy, sizeof(double));
    if (s->n_events) {
        s->ev = (event_t*)calloc(s->n_events, sizeof(event_t));
        s->active = (size_t*)calloc(s->n_events, sizeof(size_t));
        for (size_t i = 0; i < s->n_events; i++) {
            s->ev[i].fn = p->event_fns[i];
            s->ev[i].max_allowed = p->event_max[i];
            s->ev[i].direction = p->event_dirs[i];
            s->ev[i].last_root = NAN;
            s->ev[i].y_at_root = (double*)calloc(s->ndy, sizeof(double));
        }
    }
    /* K zero-filled, rk.cpp:151-157 */
    s->K = (double*)calloc(s->ny * (size_t)s->tab->K_stride, sizeof(double));

    s->t_tmp = 0.0; s->t_now = s->t_start; s->t_old = s->t_start; s->len_t = 0;
    memcpy(s->y_old, p->y0, sizeof(double) * s->ny);
    memcpy(s->y_now, p->y0, sizeof(double) * s->ny);
    call_diffeq(s);                                         /* RHS evaluation #1, cysolver.cpp:475 */
    memcpy(s->dy_old, s->dy_now, sizeof(double) * s->ndy);

    if (s->n_events) {                                      /* cysolver.cpp:481-510 */
        double* y_use = s->y_now;
        if (s->capture_extra) {
            memcpy(s->dy_tmp2, s->y_now, sizeof(double) * s->ny);
            memcpy(s->dy_tmp2 + s->ny, s->dy_now + s->ny, sizeof(double) * s->nextra);
            y_use = s->dy_tmp2;
        }
        for (size_t i = 0; i < s->n_events; i++) s->ev[i].g_old = s->ev[i].fn(s->t_now, y_use, s->args);
    }
    if (!s->use_t_eval) save_data(s, s->t_now, s->y_now, s->dy_now);   /* :513-517 */
    s->t_eval_index_old = s->fwd ? 0 : s->n_t_eval;                    /* :520-527 */

    /* Tolerances must be known before the t0 interpolant?  No: build_dense(true) at :530-533 only
       needs K (zeros) and the state; it runs BEFORE the RK-specific part of setup. */
    s->n_rtols = p->n_rtols; s->n_atols = p->n_atols;
    if (s->dense_output) build_dense(s, 1);

    /* --- RKSolver::setup, rk.cpp:560-618 */
    s->first_step = p->first_step;
    s->max_step_size = p->max_step;
    if (s->first_step != 0.0) {
        /* The reference marks -80 here but then overwrites it / crashes (quirk 10);
           the restatement returns -80 cleanly, as the CUDA build must. */
        if (s->first_step < 0.0 || s->first_step > (s->t_delta_abs * 0.5)) { s->status = ST_BAD_FIRST_STEP; return s; }
    }
    if (s->n_rtols == 0 || (s->n_rtols > 1 && s->n_rtols != s->ny) || s->n_atols == 0 || (s->n_atols > 1 && s->n_atols != s->ny)) {
        s->status = ST_BAD_CONFIG; return s;
    }
    s->arr_rtol = s->n_rtols > 1; s->arr_atol = s->n_atols > 1;
    s->rtols = (double*)malloc(sizeof(double) * s->n_rtols);
    s->atols = (double*)malloc(sizeof(double) * s->n_atols);
    memcpy(s->atols, p->atols, sizeof(double) * s->n_atols);
    for (size_t i = 0; i < s->n_rtols; i++) {               /* rtol floor, rk.cpp:600-608 */
        double r = p->rtols[i];
        if (r < EPS_100) r = EPS_100;
        s->rtols[i] = r;
    }
    if (s->first_step == 0.0) calc_first_step_size(s);      /* RHS evaluation #2 */
    else s->step_size = s->first_step;

    /* --- back in CySolverResult::setup, cysolution.cpp:375-382: the single re-used interpolant */
    if ((s->use_t_eval && !s->dense_output) || (s->n_events > 0 && !s->dense_output)) {
        s->dense = (dense_t*)realloc(s->dense, sizeof(dense_t) * 1);
        s->dense_cap = 1; s->n_dense = 1;
        s->dense[0].state = (double*)calloc(s->ny * (s->tab->Q_order + 1), sizeof(double));
        s->dense[0].t_old = 0.0; s->dense[0].t_now = 0.0; s->dense[0].step = 0.0;
        s->num_interpolates++;
    }
    s->setup_ok = 1;
    s->status = ST_NO_ERROR;
    return s;
}

static solver_t* solver_run(const orc_problem* p)
{
    solver_t* s = solver_setup(p);
    /* cy/cysolver.cpp:994-1000 with check_status() :550-561 */
    while (s->status == ST_NO_ERROR && !s->error_flag && s->setup_ok) take_step(s);
    if (s->status == ST_SUCCESS || s->status == ST_EVENT_TERMINATED) s->success = 1;
    return s;
}

/* ---- post-hoc dense evaluation: cy/cysolution.cpp:511-596 */
static int result_call(solver_t* s, double t, double* out)
{
    if (!s->success) return ST_NOT_SUCCESSFUL;
    if (!s->dense_output) return ST_DENSE_NOT_SAVED;
    if (s->n_dense == 0) return ST_DENSE_NOT_SAVED;
    if (!out) return ST_ARGUMENT_ERROR;
    const double tmin = fmin(s->t_start, s->t_end), tmax = fmax(s->t_start, s->t_end);
    if ((t < tmin) || (t > tmax)) return ST_ARGUMENT_ERROR;
    const size_t len = s->use_t_eval ? s->num_interpolates : s->size;
    const double* times = s->use_t_eval ? s->interp_time : s->t_out;
    size_t idx = 0;
    if (s->n_dense > 1) {
        if (s->fwd) {
            idx = upper_bound_d(times, len, t);
        } else {
            /* `times` is descending; the reference searches its ascending mirror
               (p_finalize, :159-172): lower_bound there == count of elements < t. */
            size_t lo = 0;
            for (size_t i = 0; i < len; i++) if (times[len - 1 - i] < t) lo = i + 1; else break;
            idx = len - lo - 1;
        }
    }
    if (idx >= s->n_dense) idx = s->n_dense - 1;
    dense_call(s, &s->dense[idx], t, out);
    return ST_NO_ERROR;
}

/* =========================================================================
 * Flat interface (oracle_api.h)
 * ========================================================================= */
void*  orc_solve(const orc_problem* p) { return solver_run(p); }
void   orc_free(void* h) { solver_free((solver_t*)h); }
#define S(h) ((solver_t*)(h))
int    orc_status(void* h) { return S(h)->status; }
int    orc_success(void* h) { return S(h)->success; }
size_t orc_size(void* h) { return S(h)->size; }
size_t orc_steps_taken(void* h) { return S(h)->steps_taken; }
size_t orc_num_interpolates(void* h) { return S(h)->num_interpolates; }
size_t orc_num_dy(void* h) { return S(h)->ndy; }
size_t orc_nfev(void* h) { return S(h)->nfev; }
const double* orc_t_ptr(void* h) { return S(h)->t_out; }
const double* orc_y_ptr(void* h) { return S(h)->y_out; }
int    orc_event_terminated(void* h) { return S(h)->event_terminated; }
size_t orc_event_terminate_index(void* h) { return S(h)->event_terminate_index; }
size_t orc_event_count(void* h, size_t e) { return e < S(h)->n_events ? S(h)->ev[e].n_hits : 0; }
const double* orc_event_t_ptr(void* h, size_t e) { return S(h)->ev[e].t_hits; }
const double* orc_event_y_ptr(void* h, size_t e) { return S(h)->ev[e].y_hits; }

int orc_call(void* h, const double* t, size_t n, double* out)   /* call_vectorize, cysolution.cpp:598-616 */
{
    int st = ST_NO_ERROR;
    for (size_t i = 0; i < n; i++) {
        st = result_call(S(h), t[i], out + S(h)->ndy * i);
        if (st != ST_NO_ERROR) break;
    }
    return st;
}

static double wall_now(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double orc_solve_batch(const orc_problem* base, size_t B, const double* y0s,
                       const char* argss, size_t args_stride,
                       const double* t_spans, int n_threads,
                       int* status, size_t* size, size_t* steps, size_t* nfev,
                       double* t_last, double* y_last, double* teval_out)
{
    const size_t ny = base->num_y, ndy = base->num_y + base->num_extra;
    if (n_threads < 1) n_threads = 1;
    const double t0 = wall_now();
#pragma omp parallel for schedule(static) num_threads(n_threads)
    for (long long i = 0; i < (long long)B; i++) {
        orc_problem p = *base;
        p.y0 = y0s + (size_t)i * ny;
        if (args_stride) p.args = argss + (size_t)i * args_stride;
        if (t_spans) { p.t_start = t_spans[2 * i]; p.t_end = t_spans[2 * i + 1]; }
        solver_t* s = solver_run(&p);
        if (status) status[i] = s->status;
        if (size) size[i] = s->size;
        if (steps) steps[i] = s->steps_taken;
        if (nfev) nfev[i] = s->nfev;
        if (s->size > 0) {
            if (t_last) t_last[i] = s->t_out[s->size - 1];
            if (y_last) memcpy(y_last + (size_t)i * ndy, s->y_out + (s->size - 1) * ndy, sizeof(double) * ndy);
            if (teval_out && base->n_t_eval) {
                const size_t rows = s->size < base->n_t_eval ? s->size : base->n_t_eval;
                memcpy(teval_out + (size_t)i * base->n_t_eval * ndy, s->y_out, sizeof(double) * rows * ndy);
            }
        }
        solver_free(s);
    }
    return wall_now() - t0;
}

static double poly3_eval(void* ctx, double x, int buf)
{
    (void)buf;
    const double* c = (const double*)ctx;
    return ((c[3] * x + c[2]) * x + c[1]) * x + c[0];
}

double orc_brentq_poly3(const double a[4], double xa, double xb, double xtol, double rtol,
                        size_t mitr, int* err, size_t* funcalls, size_t* iterations)
{
    brent_info info = { 0, 0, 0, 0 };
    double c[4] = { a[0], a[1], a[2], a[3] };
    const double r = brentq(poly3_eval, c, xa, xb, xtol, rtol, mitr, &info);
    if (err) *err = info.error_num;
    if (funcalls) *funcalls = info.funcalls;
    if (iterations) *iterations = info.iterations;
    return r;
}
