// tableaux.cuh -- Butcher tableaux of the three embedded pairs as compile-time tables.
//
// Every coefficient is a `constexpr` looked up with template indices, so each use becomes a
// literal operand served from the kernel's constant bank (c[0x2][...] in SASS) and every zero
// entry disappears at compile time -- the same entries the reference skips by hand
// (CyRK/cy/rk.cpp:738,1006,1204,1241,1390).  Term order inside a row is ascending stage index,
// which is the order the reference sums in; with -fmad=false the sums are bit-identical to it.
//
// Values: Bogacki-Shampine 3(2)   rk.cpp:695-862
//         Dormand-Prince 5(4)     rk.cpp:886-1131
//         DOP853 (Hairer)         rk.cpp:1154-1676
#pragma once
#include "cyrk_defs.h"

#define CYRK_HD __host__ __device__

namespace cyrk {

template <int I> struct IC { static constexpr int value = I; };

template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(f);
    }
}

// ---------------------------------------------------------------------------------------------
struct TabRK23 {
    static constexpr int METHOD = METHOD_RK23;
    static constexpr int NS = 3;          // stages; k[NS] is the FSAL derivative at the new point
    static constexpr int NK = NS + 1;
    static constexpr int NK_DENSE = NK;   // no extra stages for the interpolant
    static constexpr int Q_ORDER = 3;
    static constexpr bool IS_DOP = false;
    static constexpr bool STAGE1_STEP_FIRST = true;   // y_old + (step*k0*a), rk.cpp:721
    CYRK_HD static constexpr double err_exponent() { return 1.0 / (2.0 + 1.0); }
    CYRK_HD static constexpr double c(int s) { constexpr double T[NS] = {0.0, 1.0 / 2.0, 3.0 / 4.0}; return T[s]; }
    CYRK_HD static constexpr double a(int s, int j)
    {
        constexpr double T[NS][NS] = {{0, 0, 0}, {1.0 / 2.0, 0, 0}, {0, 3.0 / 4.0, 0}};
        return T[s][j];
    }
    CYRK_HD static constexpr double b(int j) { constexpr double T[NS] = {2.0 / 9.0, 1.0 / 3.0, 4.0 / 9.0}; return T[j]; }
    CYRK_HD static constexpr double e(int j) { constexpr double T[NK] = {5.0 / 72.0, -1.0 / 12.0, -1.0 / 9.0, 1.0 / 8.0}; return T[j]; }
    // interpolant polynomial, column p (p = 0 is plain k0)
    CYRK_HD static constexpr double p(int col, int j)
    {
        constexpr double T[Q_ORDER][NK] = {{1.0, 0, 0, 0},
                                           {-4.0 / 3.0, 1.0, 4.0 / 3.0, -1.0},
                                           {5.0 / 9.0, -2.0 / 3.0, -8.0 / 9.0, 1.0}};
        return T[col][j];
    }
};

struct TabRK45 {
    static constexpr int METHOD = METHOD_RK45;
    static constexpr int NS = 6;
    static constexpr int NK = NS + 1;
    static constexpr int NK_DENSE = NK;
    static constexpr int Q_ORDER = 4;
    static constexpr bool IS_DOP = false;
    static constexpr bool STAGE1_STEP_FIRST = false;  // y_old + step*(a*k0), rk.cpp:908
    CYRK_HD static constexpr double err_exponent() { return 1.0 / (4.0 + 1.0); }
    CYRK_HD static constexpr double c(int s)
    {
        constexpr double T[NS] = {0.0, 1.0 / 5.0, 3.0 / 10.0, 4.0 / 5.0, 8.0 / 9.0, 1.0};
        return T[s];
    }
    CYRK_HD static constexpr double a(int s, int j)
    {
        constexpr double T[NS][NS - 1] = {
            {0, 0, 0, 0, 0},
            {1.0 / 5.0, 0, 0, 0, 0},
            {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0},
            {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0},
            {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0},
            {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0}};
        return j < NS - 1 ? T[s][j] : 0.0;
    }
    CYRK_HD static constexpr double b(int j)
    {
        constexpr double T[NS] = {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0};
        return T[j];
    }
    CYRK_HD static constexpr double e(int j)
    {
        constexpr double T[NK] = {-71.0 / 57600.0, 0.0, 71.0 / 16695.0, -71.0 / 1920.0,
                                  17253.0 / 339200.0, -22.0 / 525.0, 1.0 / 40.0};
        return T[j];
    }
    CYRK_HD static constexpr double p(int col, int j)
    {
        constexpr double T[Q_ORDER][NK] = {
            {1.0, 0, 0, 0, 0, 0, 0},
            {-8048581381.0 / 2820520608.0, 0.0, 131558114200.0 / 32700410799.0, -1754552775.0 / 470086768.0,
             127303824393.0 / 49829197408.0, -282668133.0 / 205662961.0, 40617522.0 / 29380423.0},
            {8663915743.0 / 2820520608.0, 0.0, -68118460800.0 / 10900136933.0, 14199869525.0 / 1410260304.0,
             -318862633887.0 / 49829197408.0, 2019193451.0 / 616988883.0, -110615467.0 / 29380423.0},
            {-12715105075.0 / 11282082432.0, 0.0, 87487479700.0 / 32700410799.0, -10690763975.0 / 1880347072.0,
             701980252875.0 / 199316789632.0, -1453857185.0 / 822651844.0, 69997945.0 / 29380423.0}};
        return T[col][j];
    }
};

struct TabDOP853 {
    static constexpr int METHOD = METHOD_DOP853;
    static constexpr int NS = 12;
    static constexpr int NK = NS + 1;
    static constexpr int NK_DENSE = NK + 3;   // three extra stages feed the 7th-order interpolant
    static constexpr int Q_ORDER = 7;
    static constexpr bool IS_DOP = true;
    static constexpr bool STAGE1_STEP_FIRST = false;  // rk.cpp:1174
    CYRK_HD static constexpr double err_exponent() { return 1.0 / (7.0 + 1.0); }
    CYRK_HD static constexpr double c(int s)
    {
        constexpr double T[NS] = {0.0,
                                  0.526001519587677318785587544488e-01,
                                  0.789002279381515978178381316732e-01,
                                  0.118350341907227396726757197510,
                                  0.281649658092772603273242802490,
                                  0.333333333333333333333333333333,
                                  0.25,
                                  0.307692307692307692307692307692,
                                  0.651282051282051282051282051282,
                                  0.6,
                                  0.857142857142857142857142857142,
                                  1.0};
        return T[s];
    }
    CYRK_HD static constexpr double a(int s, int j)
    {
        constexpr double T[NS][NS - 1] = {
            {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
            {5.26001519587677318785587544488e-2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
            {1.97250569845378994544595329183e-2, 5.91751709536136983633785987549e-2, 0, 0, 0, 0, 0, 0, 0, 0, 0},
            {2.95875854768068491816892993775e-2, 0, 8.87627564304205475450678981324e-2, 0, 0, 0, 0, 0, 0, 0, 0},
            {2.41365134159266685502369798665e-1, 0, -8.84549479328286085344864962717e-1,
             9.24834003261792003115737966543e-1, 0, 0, 0, 0, 0, 0, 0},
            {3.7037037037037037037037037037e-2, 0, 0, 1.70828608729473871279604482173e-1,
             1.25467687566822425016691814123e-1, 0, 0, 0, 0, 0, 0},
            {3.7109375e-2, 0, 0, 1.70252211019544039314978060272e-1, 6.02165389804559606850219397283e-2,
             -1.7578125e-2, 0, 0, 0, 0, 0},
            {3.70920001185047927108779319836e-2, 0, 0, 1.70383925712239993810214054705e-1,
             1.07262030446373284651809199168e-1, -1.53194377486244017527936158236e-2,
             8.27378916381402288758473766002e-3, 0, 0, 0, 0},
            {6.24110958716075717114429577812e-1, 0, 0, -3.36089262944694129406857109825,
             -8.68219346841726006818189891453e-1, 2.75920996994467083049415600797e1,
             2.01540675504778934086186788979e1, -4.34898841810699588477366255144e1, 0, 0, 0},
            {4.77662536438264365890433908527e-1, 0, 0, -2.48811461997166764192642586468,
             -5.90290826836842996371446475743e-1, 2.12300514481811942347288949897e1,
             1.52792336328824235832596922938e1, -3.32882109689848629194453265587e1,
             -2.03312017085086261358222928593e-2, 0, 0},
            {-9.3714243008598732571704021658e-1, 0, 0, 5.18637242884406370830023853209,
             1.09143734899672957818500254654, -8.14978701074692612513997267357,
             -1.85200656599969598641566180701e1, 2.27394870993505042818970056734e1,
             2.49360555267965238987089396762, -3.0467644718982195003823669022, 0},
            {2.27331014751653820792359768449, 0, 0, -1.05344954667372501984066689879e1,
             -2.00087205822486249909675718444, -1.79589318631187989172765950534e1,
             2.79488845294199600508499808837e1, -2.85899827713502369474065508674,
             -8.87285693353062954433549289258, 1.23605671757943030647266201528e1,
             6.43392746015763530355970484046e-1}};
        return j < NS - 1 ? T[s][j] : 0.0;
    }
    CYRK_HD static constexpr double b(int j)
    {
        constexpr double T[NS] = {5.42937341165687622380535766363e-2, 0, 0, 0, 0,
                                  4.45031289275240888144113950566, 1.89151789931450038304281599044,
                                  -5.8012039600105847814672114227, 3.1116436695781989440891606237e-1,
                                  -1.52160949662516078556178806805e-1, 2.01365400804030348374776537501e-1,
                                  4.47106157277725905176885569043e-2};
        return T[j];
    }
    // 3rd-order error weights: B minus the bhh terms, folded in binary64 like rk.cpp:1450-1457
    CYRK_HD static constexpr double e3(int j)
    {
        return j == 0    ? (5.42937341165687622380535766363e-2 - 0.244094488188976377952755905512)
               : j == 8  ? (3.1116436695781989440891606237e-1 - 0.733846688281611857341361741547)
               : j == 11 ? (4.47106157277725905176885569043e-2 - 0.220588235294117647058823529412e-1)
               : j < NS  ? b(j)
                         : 0.0;
    }
    CYRK_HD static constexpr double e5(int j)
    {
        constexpr double T[NS] = {0.1312004499419488073250102996e-1, 0, 0, 0, 0,
                                  -0.1225156446376204440720569753e+1, -0.4957589496572501915214079952,
                                  0.1664377182454986536961530415e+1, -0.3503288487499736816886487290,
                                  0.3341791187130174790297318841, 0.8192320648511571246570742613e-1,
                                  -0.2235530786388629525884427845e-1};
        return j < NS ? T[j] : 0.0;
    }
    // extra stages (rows 13..15 of A, rk.cpp:1506-1583); column index = k index (0..14)
    CYRK_HD static constexpr double cx(int r) { constexpr double T[3] = {0.1, 0.2, 0.777777777777777777777777777778}; return T[r]; }
    CYRK_HD static constexpr double ax(int r, int j)
    {
        constexpr double T[3][15] = {
            {5.61675022830479523392909219681e-2, 0, 0, 0, 0, 0, 2.53500210216624811088794765333e-1,
             -2.46239037470802489917441475441e-1, -1.24191423263816360469010140626e-1,
             1.5329179827876569731206322685e-1, 8.20105229563468988491666602057e-3,
             7.56789766054569976138603589584e-3, -8.298e-3, 0, 0},
            {3.18346481635021405060768473261e-2, 0, 0, 0, 0, 2.83009096723667755288322961402e-2,
             5.35419883074385676223797384372e-2, -5.49237485713909884646569340306e-2, 0, 0,
             -1.08347328697249322858509316994e-4, 3.82571090835658412954920192323e-4,
             -3.40465008687404560802977114492e-4, 1.41312443674632500278074618366e-1, 0},
            {-4.28896301583791923408573538692e-1, 0, 0, 0, 0, -4.69762141536116384314449447206,
             7.68342119606259904184240953878, 4.06898981839711007970213554331,
             3.56727187455281109270669543021e-1, 0, 0, 0, -1.39902416515901462129418009734e-3,
             2.9475147891527723389556272149, -9.15095847217987001081870187138}};
        return T[r][j];
    }
    // interpolation weights D (4 rows over k0..k15), rk.cpp:1600-1657
    CYRK_HD static constexpr double d(int r, int j)
    {
        constexpr double T[4][16] = {
            {-0.84289382761090128651353491142e+1, 0, 0, 0, 0, 0.56671495351937776962531783590,
             -0.30689499459498916912797304727e+1, 0.23846676565120698287728149680e+1,
             0.21170345824450282767155149946e+1, -0.87139158377797299206789907490,
             0.22404374302607882758541771650e+1, 0.63157877876946881815570249290,
             -0.88990336451333310820698117400e-1, 0.18148505520854727256656404962e+2,
             -0.91946323924783554000451984436e+1, -0.44360363875948939664310572000e+1},
            {0.10427508642579134603413151009e+2, 0, 0, 0, 0, 0.24228349177525818288430175319e+3,
             0.16520045171727028198505394887e+3, -0.37454675472269020279518312152e+3,
             -0.22113666853125306036270938578e+2, 0.77334326684722638389603898808e+1,
             -0.30674084731089398182061213626e+2, -0.93321305264302278729567221706e+1,
             0.15697238121770843886131091075e+2, -0.31139403219565177677282850411e+2,
             -0.93529243588444783865713862664e+1, 0.35816841486394083752465898540e+2},
            {0.19985053242002433820987653617e+2, 0, 0, 0, 0, -0.38703730874935176555105901742e+3,
             -0.18917813819516756882830838328e+3, 0.52780815920542364900561016686e+3,
             -0.11573902539959630126141871134e+2, 0.68812326946963000169666922661e+1,
             -0.10006050966910838403183860980e+1, 0.77771377980534432092869265740,
             -0.27782057523535084065932004339e+1, -0.60196695231264120758267380846e+2,
             0.84320405506677161018159903784e+2, 0.11992291136182789328035130030e+2},
            {-0.25693933462703749003312586129e+2, 0, 0, 0, 0, -0.15418974869023643374053993627e+3,
             -0.23152937917604549567536039109e+3, 0.35763911791061412378285349910e+3,
             0.93405324183624310003907691704e+2, -0.37458323136451633156875139351e+2,
             0.10409964950896230045147246184e+3, 0.29840293426660503123344363579e+2,
             -0.43533456590011143754432175058e+2, 0.96324553959188282948394950600e+2,
             -0.39177261675615439165231486172e+2, -0.14972683625798562581422125276e+3}};
        return T[r][j];
    }
};

// ---------------------------------------------------------------------------------------------
// The same numbers in __constant__ memory.  The constexpr tables above decide AT COMPILE TIME which
// terms exist; the values themselves are read as constant-bank operands (DFMA R, R, c[0x3][off], R).
// Without this, sm_100a ptxas materialises every 64-bit literal with two UMOV instructions per use
// (17 % of all issued instructions in the first profile, profiles/r01_*).
// ---------------------------------------------------------------------------------------------
struct StoreRK23 { double c[3]; double a[3][3]; double b[3]; double e[4]; double p[3][4]; };
struct StoreRK45 { double c[6]; double a[6][6]; double b[6]; double e[7]; double p[4][7]; };
struct StoreDOP853 { double c[12]; double a[12][12]; double b[12]; double e3[12]; double e5[12];
                     double cx[3]; double ax[3][15]; double d[4][16]; };

CYRK_HD constexpr StoreRK23 make_store_rk23()
{
    StoreRK23 s{};
    for (int i = 0; i < 3; i++) { s.c[i] = TabRK23::c(i); s.b[i] = TabRK23::b(i); for (int j = 0; j < 3; j++) s.a[i][j] = TabRK23::a(i, j); }
    for (int j = 0; j < 4; j++) { s.e[j] = TabRK23::e(j); for (int q = 0; q < 3; q++) s.p[q][j] = TabRK23::p(q, j); }
    return s;
}
CYRK_HD constexpr StoreRK45 make_store_rk45()
{
    StoreRK45 s{};
    for (int i = 0; i < 6; i++) { s.c[i] = TabRK45::c(i); s.b[i] = TabRK45::b(i); for (int j = 0; j < 6; j++) s.a[i][j] = TabRK45::a(i, j); }
    for (int j = 0; j < 7; j++) { s.e[j] = TabRK45::e(j); for (int q = 0; q < 4; q++) s.p[q][j] = TabRK45::p(q, j); }
    return s;
}
CYRK_HD constexpr StoreDOP853 make_store_dop853()
{
    StoreDOP853 s{};
    for (int i = 0; i < 12; i++) {
        s.c[i] = TabDOP853::c(i); s.b[i] = TabDOP853::b(i); s.e3[i] = TabDOP853::e3(i); s.e5[i] = TabDOP853::e5(i);
        for (int j = 0; j < 12; j++) s.a[i][j] = TabDOP853::a(i, j);
    }
    for (int r = 0; r < 3; r++) { s.cx[r] = TabDOP853::cx(r); for (int j = 0; j < 15; j++) s.ax[r][j] = TabDOP853::ax(r, j); }
    for (int r = 0; r < 4; r++) for (int j = 0; j < 16; j++) s.d[r][j] = TabDOP853::d(r, j);
    return s;
}

static __constant__ StoreRK23 g_store_rk23 = make_store_rk23();
static __constant__ StoreRK45 g_store_rk45 = make_store_rk45();
static __constant__ StoreDOP853 g_store_dop853 = make_store_dop853();


// ---------------------------------------------------------------------------------------------
// Step-size controller constants of the production build, also in __constant__ memory.
// With e2 = NY * err^2 (the un-normalised squared error norm) the reference's
//     SAFETY * err^(-1/(q+1))  =  0.9 * NY^(1/M) * e2^(-1/M),   M = 2(q+1) = 6, 10, 16
// and one Newton step r1 = r*((M+1) - e2*r^M)/M from a seed r ~ e2^(-1/M) gives
//     0.9 * err^(-1/(q+1))  ~=  r * (A - B * e2 * r^M),   A = C(M+1)/M,  B = C/M,  C = 0.9 * NY^(1/M).
// ---------------------------------------------------------------------------------------------
CYRK_HD constexpr double const_root(double x, int m)      // x^(1/m), x >= 1, by Newton's method
{
    double r = 1.0 + (x - 1.0) / m;
    for (int it = 0; it < 200; it++) {
        double pw = 1.0;
        for (int i = 0; i < m - 1; i++) pw *= r;
        r = r - (r * pw - x) / (m * pw);
    }
    return r;
}
struct StoreCtrl { double A[3][CYRK_MAX_NY + 1]; double B[3][CYRK_MAX_NY + 1]; double thresh_rel, thresh_abs; };
CYRK_HD constexpr StoreCtrl make_store_ctrl()
{
    StoreCtrl s{};
    const int ms[3] = {6, 10, 16};
    for (int mi = 0; mi < 3; mi++)
        for (int n = 1; n <= CYRK_MAX_NY; n++) {
            const double C = 0.9 * const_root((double)n, ms[mi]);
            s.A[mi][n] = C * (double)(ms[mi] + 1) / (double)ms[mi];
            s.B[mi][n] = C / (double)ms[mi];
        }
    s.thresh_rel = 4.0e-15;
    s.thresh_abs = 1.0e-300;
    return s;
}
static __constant__ StoreCtrl g_store_ctrl = make_store_ctrl();

template <class Tab> struct TabLd;
template <> struct TabLd<TabRK23> {
    __device__ __forceinline__ static double c(int s) { return g_store_rk23.c[s]; }
    __device__ __forceinline__ static double a(int s, int j) { return g_store_rk23.a[s][j]; }
    __device__ __forceinline__ static double b(int j) { return g_store_rk23.b[j]; }
    __device__ __forceinline__ static double e(int j) { return g_store_rk23.e[j]; }
    __device__ __forceinline__ static double p(int q, int j) { return g_store_rk23.p[q][j]; }
};
template <> struct TabLd<TabRK45> {
    __device__ __forceinline__ static double c(int s) { return g_store_rk45.c[s]; }
    __device__ __forceinline__ static double a(int s, int j) { return g_store_rk45.a[s][j]; }
    __device__ __forceinline__ static double b(int j) { return g_store_rk45.b[j]; }
    __device__ __forceinline__ static double e(int j) { return g_store_rk45.e[j]; }
    __device__ __forceinline__ static double p(int q, int j) { return g_store_rk45.p[q][j]; }
};
template <> struct TabLd<TabDOP853> {
    __device__ __forceinline__ static double c(int s) { return g_store_dop853.c[s]; }
    __device__ __forceinline__ static double a(int s, int j) { return g_store_dop853.a[s][j]; }
    __device__ __forceinline__ static double b(int j) { return g_store_dop853.b[j]; }
    __device__ __forceinline__ static double e3(int j) { return g_store_dop853.e3[j]; }
    __device__ __forceinline__ static double e5(int j) { return g_store_dop853.e5[j]; }
    __device__ __forceinline__ static double cx(int r) { return g_store_dop853.cx[r]; }
    __device__ __forceinline__ static double ax(int r, int j) { return g_store_dop853.ax[r][j]; }
    __device__ __forceinline__ static double d(int r, int j) { return g_store_dop853.d[r][j]; }
    // unused by this method; present so generic code compiles
    __device__ __forceinline__ static double e(int) { return 0.0; }
    __device__ __forceinline__ static double p(int, int) { return 0.0; }
};

template <int METHOD> struct TabFor;
template <> struct TabFor<METHOD_RK23> { using type = TabRK23; };
template <> struct TabFor<METHOD_RK45> { using type = TabRK45; };
template <> struct TabFor<METHOD_DOP853> { using type = TabDOP853; };

}  // namespace cyrk
