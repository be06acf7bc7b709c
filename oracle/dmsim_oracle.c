/* oracle/dmsim_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the
 * product path (dm-sim_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker.
 *
 * Plain-C restatement of pnnl/DM-Sim's density-matrix gate-application algorithm
 * ("formula transformation": forward pass over columns, conjugate transpose, backward pass), following
 * the reference file src/dmsim_nvgpu_omp.cuh (GPU semantics) / src/dmsim_cpu_omp.hpp (same arithmetic).
 * Each function cites the reference lines it restates.  Arithmetic is written in the reference's
 * evaluation order so that, compiled without FMA contraction, it is bit-identical to the reference
 * CPU backend (pinned by tests/test_oracle.py against oracle/_ref and tests/golden/).
 *
 * Deliberate difference: the adjoint uses the GPU backend's (correct) tile transpose
 * (src/dmsim_nvgpu_omp.cuh:825-855); the CPU backend's serialised version (src/dmsim_cpu_omp.hpp:791-825)
 * reads its scratch tile before it is filled and is only right for one sim() from the reset state.
 * Consequently repeated orc_sim() calls follow the GPU semantics (state continues correctly).
 *
 * Layout (src/dmsim_nvgpu_omp.cuh:989-998, :230-240): split FP64 arrays re[], im[] of 4^n entries,
 * flat index = col*dim + row, qubit 0 = least-significant bit of row.  What is stored is rho^T.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.14159265358979323846  /* src/config.hpp:55 */
#define ORC_S2I 0.70710678118654752440 /* src/config.hpp:57 */

typedef uint64_t idx_t;

/* enum OP order, src/dmsim_nvgpu_omp.cuh:42-48 */
enum
{
    OP_U3, OP_U2, OP_U1, OP_CX, OP_ID, OP_X, OP_Y, OP_Z, OP_H, OP_S,
    OP_SDG, OP_T, OP_TDG, OP_RX, OP_RY, OP_RZ, OP_CZ, OP_CY, OP_SWAP, OP_CH,
    OP_CCX, OP_CSWAP, OP_CRX, OP_CRY, OP_CRZ, OP_CU1, OP_CU3, OP_RXX, OP_RZZ, OP_RCCX,
    OP_RC3X, OP_C3X, OP_C3SQRTX, OP_C4X, OP_R, OP_SRN, OP_W, OP_RYY,
    OP_RAW_C1 = 100, OP_RAW_C2 = 101
};

/* POD mirror of class Gate, src/dmsim_nvgpu_omp.cuh:99-191 (no op pointer) */
typedef struct
{
    int32_t op;
    int32_t qb[5];
    double theta, phi, lambda;
    int64_t mat; /* raw C1/C2: index of a 32-double (re,im interleaved, row-major) matrix */
} orc_gate;

typedef struct
{
    int n;
    idx_t dim;
    double *re, *im;     /* current state (result of the last sim) */
    double *bre, *bim;   /* second buffer (dm_*_buf) */
} orc_sim_t;

/* ---- one pass context: the arrays a pass works on ---- */
typedef struct
{
    int n;
    idx_t dim;
    double *re, *im;
} pass_t;

/* OP_HEAD, src/dmsim_nvgpu_omp.cuh:989-998: pair enumeration for a 1-qubit op on row bit `qubit` */
#define OP_HEAD                                                          \
    const idx_t half_dim = p->dim >> 1;                                  \
    const idx_t total = half_dim * p->dim;                               \
    double *dm_real = p->re, *dm_imag = p->im;                           \
    _Pragma("omp parallel for schedule(static)")                         \
    for (idx_t i = 0; i < total; i++)                                    \
    {                                                                    \
        idx_t col = (i >> (p->n - 1));                                   \
        idx_t outer = ((i & (half_dim - 1)) >> qubit);                   \
        idx_t inner = (i & (((idx_t)1 << qubit) - 1));                   \
        idx_t offset = (outer << (qubit + 1));                           \
        idx_t pos0 = (col << p->n) + offset + inner;                     \
        idx_t pos1 = pos0 + ((idx_t)1 << qubit);
#define OP_TAIL }

/* C1_GATE, src/dmsim_nvgpu_omp.cuh:1004-1025 */
static void C1_GATE(pass_t* p, double e0_real, double e0_imag, double e1_real, double e1_imag,
                    double e2_real, double e2_imag, double e3_real, double e3_imag, int qubit)
{
    OP_HEAD
    const double el0_real = dm_real[pos0];
    const double el0_imag = dm_imag[pos0];
    const double el1_real = dm_real[pos1];
    const double el1_imag = dm_imag[pos1];
    dm_real[pos0] = (e0_real * el0_real) - (e0_imag * el0_imag) + (e1_real * el1_real) - (e1_imag * el1_imag);
    dm_imag[pos0] = (e0_real * el0_imag) + (e0_imag * el0_real) + (e1_real * el1_imag) + (e1_imag * el1_real);
    dm_real[pos1] = (e2_real * el0_real) - (e2_imag * el0_imag) + (e3_real * el1_real) - (e3_imag * el1_imag);
    dm_imag[pos1] = (e2_real * el0_imag) + (e2_imag * el0_real) + (e3_real * el1_imag) + (e3_imag * el1_real);
    OP_TAIL
}

/* index decomposition shared by C2_GATE / CX_GATE, src/dmsim_nvgpu_omp.cuh:1049-1069, :1137-1155 */
typedef struct
{
    idx_t q0dim, q1dim, outer_factor, mider_factor, inner_factor, per_col;
} quad_t;

static quad_t quad_setup(const pass_t* p, int qa, int qb)
{
    quad_t q;
    int mx = qa > qb ? qa : qb, mn = qa > qb ? qb : qa;
    q.q0dim = (idx_t)1 << mx;
    q.q1dim = (idx_t)1 << mn;
    q.outer_factor = (p->dim + q.q0dim + q.q0dim - 1) >> (mx + 1);
    q.mider_factor = (q.q0dim + q.q1dim + q.q1dim - 1) >> (mn + 1);
    q.inner_factor = q.q1dim;
    q.per_col = q.outer_factor * q.mider_factor * q.inner_factor;
    return q;
}

/* C2_GATE, src/dmsim_nvgpu_omp.cuh:1028-1122.  m = 16 complex entries, row-major, (re,im) pairs.
 * Matrix row/col index = 2*bit(qubit1) + bit(qubit2)  (pos1 = +2^qubit2, pos2 = +2^qubit1). */
static void C2_GATE(pass_t* p, const double* m, int qubit1, int qubit2)
{
    const quad_t q = quad_setup(p, qubit1, qubit2);
    const idx_t qubit1_dim = (idx_t)1 << qubit1, qubit2_dim = (idx_t)1 << qubit2;
    const idx_t total = q.per_col * p->dim;
    double *dm_real = p->re, *dm_imag = p->im;
#pragma omp parallel for schedule(static)
    for (idx_t i = 0; i < total; i++)
    {
        idx_t col = i / q.per_col;
        idx_t row = i % q.per_col;
        idx_t outer = ((row / q.inner_factor) / (q.mider_factor)) * (q.q0dim + q.q0dim);
        idx_t mider = ((row / q.inner_factor) % (q.mider_factor)) * (q.q1dim + q.q1dim);
        idx_t inner = row % q.inner_factor;
        idx_t pos0 = col * p->dim + outer + mider + inner;
        idx_t pos1 = pos0 + qubit2_dim;
        idx_t pos2 = pos0 + qubit1_dim;
        idx_t pos3 = pos0 + q.q0dim + q.q1dim;
        const idx_t pos[4] = {pos0, pos1, pos2, pos3};
        double er[4], ei[4];
        for (int k = 0; k < 4; k++) { er[k] = dm_real[pos[k]]; ei[k] = dm_imag[pos[k]]; }
        for (int r = 0; r < 4; r++)
        {
            const double* e = m + 8 * r; /* e[2c] = re, e[2c+1] = im of entry (r,c) */
            /* left-to-right sum exactly as written at :1086-1119 */
            dm_real[pos[r]] = (e[0] * er[0]) - (e[1] * ei[0]) + (e[2] * er[1]) - (e[3] * ei[1])
                            + (e[4] * er[2]) - (e[5] * ei[2]) + (e[6] * er[3]) - (e[7] * ei[3]);
            dm_imag[pos[r]] = (e[0] * ei[0]) + (e[1] * er[0]) + (e[2] * ei[1]) + (e[3] * er[1])
                            + (e[4] * ei[2]) + (e[5] * er[2]) + (e[6] * ei[3]) + (e[7] * er[3]);
        }
    }
}

/* CX_GATE, src/dmsim_nvgpu_omp.cuh:1132-1168 */
static void CX_GATE(pass_t* p, int ctrl, int qubit)
{
    const quad_t q = quad_setup(p, ctrl, qubit);
    const idx_t ctrldim = (idx_t)1 << ctrl;
    const idx_t total = q.per_col * p->dim;
    double *dm_real = p->re, *dm_imag = p->im;
#pragma omp parallel for schedule(static)
    for (idx_t i = 0; i < total; i++)
    {
        idx_t col = i / q.per_col;
        idx_t row = i % q.per_col;
        idx_t outer = ((row / q.inner_factor) / (q.mider_factor)) * (q.q0dim + q.q0dim);
        idx_t mider = ((row / q.inner_factor) % (q.mider_factor)) * (q.q1dim + q.q1dim);
        idx_t inner = row % q.inner_factor;
        idx_t pos0 = col * p->dim + outer + mider + inner + ctrldim;
        idx_t pos1 = col * p->dim + outer + mider + inner + q.q0dim + q.q1dim;
        const double el0_real = dm_real[pos0], el0_imag = dm_imag[pos0];
        const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
        dm_real[pos0] = el1_real; dm_imag[pos0] = el1_imag;
        dm_real[pos1] = el0_real; dm_imag[pos1] = el0_imag;
    }
}

/* X_GATE :1175-1187 */
static void X_GATE(pass_t* p, int qubit)
{
    OP_HEAD
    const double el0_real = dm_real[pos0], el0_imag = dm_imag[pos0];
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos0] = el1_real; dm_imag[pos0] = el1_imag;
    dm_real[pos1] = el0_real; dm_imag[pos1] = el0_imag;
    OP_TAIL
}
/* Y_GATE :1196-1209 */
static void Y_GATE(pass_t* p, int qubit)
{
    OP_HEAD
    const double el0_real = dm_real[pos0], el0_imag = dm_imag[pos0];
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos0] = el1_imag;  dm_imag[pos0] = -el1_real;
    dm_real[pos1] = -el0_imag; dm_imag[pos1] = el0_real;
    OP_TAIL
}
/* Z_GATE :1216-1225 */
static void Z_GATE(pass_t* p, int qubit)
{
    OP_HEAD
    (void)pos0;
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos1] = -el1_real; dm_imag[pos1] = -el1_imag;
    OP_TAIL
}
/* H_GATE :1232-1245 */
static void H_GATE(pass_t* p, int qubit)
{
    OP_HEAD
    const double el0_real = dm_real[pos0], el0_imag = dm_imag[pos0];
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos0] = ORC_S2I * (el0_real + el1_real);
    dm_imag[pos0] = ORC_S2I * (el0_imag + el1_imag);
    dm_real[pos1] = ORC_S2I * (el0_real - el1_real);
    dm_imag[pos1] = ORC_S2I * (el0_imag - el1_imag);
    OP_TAIL
}
/* SRN_GATE :1253-1266 (not complex-linear; restated literally) */
static void SRN_GATE(pass_t* p, int qubit)
{
    OP_HEAD
    const double el0_real = dm_real[pos0], el0_imag = dm_imag[pos0];
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos0] = 0.5 * (el0_real + el1_real);
    dm_imag[pos0] = 0.5 * (el0_imag - el1_imag);
    dm_real[pos1] = 0.5 * (el0_real + el1_real);
    dm_imag[pos1] = 0.5 * (-el0_imag + el1_imag);
    OP_TAIL
}
/* R_GATE :1283-1292 : v1 *= i*phase */
static void R_GATE(pass_t* p, double phase, int qubit)
{
    OP_HEAD
    (void)pos0;
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos1] = -(el1_imag * phase);
    dm_imag[pos1] = el1_real * phase;
    OP_TAIL
}
/* S_GATE :1299-1307 */
static void S_GATE(pass_t* p, int qubit)
{
    OP_HEAD
    (void)pos0;
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos1] = -el1_imag; dm_imag[pos1] = el1_real;
    OP_TAIL
}
/* SDG_GATE :1314-1322 */
static void SDG_GATE(pass_t* p, int qubit)
{
    OP_HEAD
    (void)pos0;
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos1] = el1_imag; dm_imag[pos1] = -el1_real;
    OP_TAIL
}
/* T_GATE :1329-1337 */
static void T_GATE(pass_t* p, int qubit)
{
    OP_HEAD
    (void)pos0;
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos1] = ORC_S2I * (el1_real - el1_imag);
    dm_imag[pos1] = ORC_S2I * (el1_real + el1_imag);
    OP_TAIL
}
/* TDG_GATE :1344-1352 */
static void TDG_GATE(pass_t* p, int qubit)
{
    OP_HEAD
    (void)pos0;
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos1] = ORC_S2I * (el1_real + el1_imag);
    dm_imag[pos1] = ORC_S2I * (-el1_real + el1_imag);
    OP_TAIL
}
/* U1_GATE :1381-1397 */
static void U1_GATE(pass_t* p, double lambda, int qubit)
{
    double e3_real = cos(lambda);
    double e3_imag = sin(lambda);
    OP_HEAD
    const double el0_real = dm_real[pos0], el0_imag = dm_imag[pos0];
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos0] = el0_real;
    dm_imag[pos0] = el0_imag;
    dm_real[pos1] = (e3_real * el1_real) - (e3_imag * el1_imag);
    dm_imag[pos1] = (e3_real * el1_imag) + (e3_imag * el1_real);
    OP_TAIL
}
/* U2_GATE :1404-1418 */
static void U2_GATE(pass_t* p, double phi, double lambda, int qubit)
{
    double e0_real = ORC_S2I, e0_imag = 0;
    double e1_real = -ORC_S2I * cos(lambda), e1_imag = -ORC_S2I * sin(lambda);
    double e2_real = ORC_S2I * cos(phi), e2_imag = ORC_S2I * sin(phi);
    double e3_real = ORC_S2I * cos(phi + lambda), e3_imag = ORC_S2I * sin(phi + lambda);
    C1_GATE(p, e0_real, e0_imag, e1_real, e1_imag, e2_real, e2_imag, e3_real, e3_imag, qubit);
}
/* U3_GATE :1425-1440 */
static void U3_GATE(pass_t* p, double theta, double phi, double lambda, int qubit)
{
    double e0_real = cos(theta / 2.), e0_imag = 0;
    double e1_real = -cos(lambda) * sin(theta / 2.), e1_imag = -sin(lambda) * sin(theta / 2.);
    double e2_real = cos(phi) * sin(theta / 2.), e2_imag = sin(phi) * sin(theta / 2.);
    double e3_real = cos(phi + lambda) * cos(theta / 2.), e3_imag = sin(phi + lambda) * cos(theta / 2.);
    C1_GATE(p, e0_real, e0_imag, e1_real, e1_imag, e2_real, e2_imag, e3_real, e3_imag, qubit);
}
/* RX_GATE :1444-1459 */
static void RX_GATE(pass_t* p, double theta, int qubit)
{
    double rx_real = cos(theta / 2.0);
    double rx_imag = -sin(theta / 2.0);
    OP_HEAD
    const double el0_real = dm_real[pos0], el0_imag = dm_imag[pos0];
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos0] = (rx_real * el0_real) - (rx_imag * el1_imag);
    dm_imag[pos0] = (rx_real * el0_imag) + (rx_imag * el1_real);
    dm_real[pos1] = -(rx_imag * el0_imag) + (rx_real * el1_real);
    dm_imag[pos1] = +(rx_imag * el0_real) + (rx_real * el1_imag);
    OP_TAIL
}
/* RY_GATE :1463-1481 */
static void RY_GATE(pass_t* p, double theta, int qubit)
{
    double e0_real = cos(theta / 2.0);
    double e1_real = -sin(theta / 2.0);
    double e2_real = sin(theta / 2.0);
    double e3_real = cos(theta / 2.0);
    OP_HEAD
    const double el0_real = dm_real[pos0], el0_imag = dm_imag[pos0];
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos0] = (e0_real * el0_real) + (e1_real * el1_real);
    dm_imag[pos0] = (e0_real * el0_imag) + (e1_real * el1_imag);
    dm_real[pos1] = (e2_real * el0_real) + (e3_real * el1_real);
    dm_imag[pos1] = (e2_real * el0_imag) + (e3_real * el1_imag);
    OP_TAIL
}
/* RZ_GATE :1485-1489  (== U1) */
static void RZ_GATE(pass_t* p, double phi, int qubit) { U1_GATE(p, phi, qubit); }
/* W_GATE :1787-1799 */
static void W_GATE(pass_t* p, int qubit)
{
    OP_HEAD
    const double el0_real = dm_real[pos0], el0_imag = dm_imag[pos0];
    const double el1_real = dm_real[pos1], el1_imag = dm_imag[pos1];
    dm_real[pos0] = ORC_S2I * (el0_real + el1_imag);
    dm_imag[pos0] = ORC_S2I * (el0_imag - el1_real);
    dm_real[pos1] = ORC_S2I * (el0_imag + el1_real);
    dm_imag[pos1] = ORC_S2I * (-el0_real + el1_imag);
    OP_TAIL
}

/* ---- composites, src/dmsim_nvgpu_omp.cuh:1493-1780, :1803-1813 (qelib1.inc order) ---- */
static void CZ_GATE(pass_t* p, int a, int b) { H_GATE(p, b); CX_GATE(p, a, b); H_GATE(p, b); }       /* :1493 */
static void CY_GATE(pass_t* p, int a, int b) { SDG_GATE(p, b); CX_GATE(p, a, b); S_GATE(p, b); }     /* :1503 */
static void CH_GATE(pass_t* p, int a, int b)                                                         /* :1513 */
{
    H_GATE(p, b); SDG_GATE(p, b); CX_GATE(p, a, b); H_GATE(p, b); T_GATE(p, b); CX_GATE(p, a, b);
    T_GATE(p, b); H_GATE(p, b); S_GATE(p, b); X_GATE(p, b); S_GATE(p, a);
}
static void CRZ_GATE(pass_t* p, double lambda, int a, int b)                                         /* :1531 */
{
    U1_GATE(p, lambda / 2, b); CX_GATE(p, a, b); U1_GATE(p, -lambda / 2, b); CX_GATE(p, a, b);
}
static void CU1_GATE(pass_t* p, double lambda, int a, int b)                                         /* :1542 */
{
    U1_GATE(p, lambda / 2, a); CX_GATE(p, a, b); U1_GATE(p, -lambda / 2, b); CX_GATE(p, a, b);
    U1_GATE(p, lambda / 2, b);
}
static void CU3_GATE(pass_t* p, double theta, double phi, double lambda, int c, int t)               /* :1554 */
{
    double temp1 = (lambda - phi) / 2;
    double temp2 = theta / 2;
    double temp3 = -(phi + lambda) / 2;
    U1_GATE(p, -temp3, c); U1_GATE(p, temp1, t); CX_GATE(p, c, t);
    U3_GATE(p, -temp2, 0, temp3, t); CX_GATE(p, c, t); U3_GATE(p, temp2, phi, 0, t);
}
static void CCX_GATE(pass_t* p, int a, int b, int c)                                                 /* :1570 */
{
    H_GATE(p, c); CX_GATE(p, b, c); TDG_GATE(p, c); CX_GATE(p, a, c); T_GATE(p, c); CX_GATE(p, b, c);
    TDG_GATE(p, c); CX_GATE(p, a, c); T_GATE(p, b); T_GATE(p, c); H_GATE(p, c); CX_GATE(p, a, b);
    T_GATE(p, a); TDG_GATE(p, b); CX_GATE(p, a, b);
}
static void SWAP_GATE(pass_t* p, int a, int b) { CX_GATE(p, a, b); CX_GATE(p, b, a); CX_GATE(p, a, b); } /* :1591 */
static void CSWAP_GATE(pass_t* p, int a, int b, int c)                                               /* :1600 */
{
    CX_GATE(p, c, b); CCX_GATE(p, a, b, c); CX_GATE(p, c, b);
}
static void CRX_GATE(pass_t* p, double lambda, int a, int b)                                         /* :1610 */
{
    U1_GATE(p, ORC_PI / 2, b); CX_GATE(p, a, b); U3_GATE(p, -lambda / 2, 0, 0, b); CX_GATE(p, a, b);
    U3_GATE(p, lambda / 2, -ORC_PI / 2, 0, b);
}
static void CRY_GATE(pass_t* p, double lambda, int a, int b)                                         /* :1622 */
{
    U3_GATE(p, lambda / 2, 0, 0, b); CX_GATE(p, a, b); U3_GATE(p, -lambda / 2, 0, 0, b); CX_GATE(p, a, b);
}
static void RXX_GATE(pass_t* p, double theta, int a, int b)                                          /* :1633 */
{
    U3_GATE(p, ORC_PI / 2, theta, 0, a); H_GATE(p, b); CX_GATE(p, a, b); U1_GATE(p, -theta, b);
    CX_GATE(p, a, b); H_GATE(p, b); U2_GATE(p, -ORC_PI, ORC_PI - theta, a);
}
static void RZZ_GATE(pass_t* p, double theta, int a, int b)                                          /* :1647 */
{
    CX_GATE(p, a, b); U1_GATE(p, theta, b); CX_GATE(p, a, b);
}
static void RCCX_GATE(pass_t* p, int a, int b, int c)                                                /* :1657 */
{
    U2_GATE(p, 0, ORC_PI, c); U1_GATE(p, ORC_PI / 4, c); CX_GATE(p, b, c); U1_GATE(p, -ORC_PI / 4, c);
    CX_GATE(p, a, c); U1_GATE(p, ORC_PI / 4, c); CX_GATE(p, b, c); U1_GATE(p, -ORC_PI / 4, c);
    U2_GATE(p, 0, ORC_PI, c);
}
static void RC3X_GATE(pass_t* p, int a, int b, int c, int d)                                         /* :1673 */
{
    U2_GATE(p, 0, ORC_PI, d); U1_GATE(p, ORC_PI / 4, d); CX_GATE(p, c, d); U1_GATE(p, -ORC_PI / 4, d);
    U2_GATE(p, 0, ORC_PI, d); CX_GATE(p, a, d); U1_GATE(p, ORC_PI / 4, d); CX_GATE(p, b, d);
    U1_GATE(p, -ORC_PI / 4, d); CX_GATE(p, a, d); U1_GATE(p, ORC_PI / 4, d); CX_GATE(p, b, d);
    U1_GATE(p, -ORC_PI / 4, d); U2_GATE(p, 0, ORC_PI, d); U1_GATE(p, ORC_PI / 4, d); CX_GATE(p, c, d);
    U1_GATE(p, -ORC_PI / 4, d); U2_GATE(p, 0, ORC_PI, d);
}
/* C3X_GATE :1698-1728 and C3SQRTX_GATE :1733-1763 share the shape, angle = PI/4 or PI/8 */
static void C3X_LIKE(pass_t* p, double ang, int a, int b, int c, int d)
{
    H_GATE(p, d); CU1_GATE(p, -ang, a, d); H_GATE(p, d);
    CX_GATE(p, a, b);
    H_GATE(p, d); CU1_GATE(p, ang, b, d); H_GATE(p, d);
    CX_GATE(p, a, b);
    H_GATE(p, d); CU1_GATE(p, -ang, b, d); H_GATE(p, d);
    CX_GATE(p, b, c);
    H_GATE(p, d); CU1_GATE(p, ang, c, d); H_GATE(p, d);
    CX_GATE(p, a, c);
    H_GATE(p, d); CU1_GATE(p, -ang, c, d); H_GATE(p, d);
    CX_GATE(p, b, c);
    H_GATE(p, d); CU1_GATE(p, ang, c, d); H_GATE(p, d);
    CX_GATE(p, a, c);
    H_GATE(p, d); CU1_GATE(p, -ang, c, d); H_GATE(p, d);
}
static void C3X_GATE(pass_t* p, int a, int b, int c, int d) { C3X_LIKE(p, ORC_PI / 4, a, b, c, d); }
static void C3SQRTX_GATE(pass_t* p, int a, int b, int c, int d) { C3X_LIKE(p, ORC_PI / 8, a, b, c, d); }
static void C4X_GATE(pass_t* p, int a, int b, int c, int d, int e)                                   /* :1767 */
{
    H_GATE(p, e); CU1_GATE(p, -ORC_PI / 2, d, e); H_GATE(p, e);
    C3X_GATE(p, a, b, c, d);
    H_GATE(p, d); CU1_GATE(p, ORC_PI / 4, d, e); H_GATE(p, d);
    C3X_GATE(p, a, b, c, d);
    C3SQRTX_GATE(p, a, b, c, e);
}
static void RYY_GATE(pass_t* p, double theta, int a, int b)                                          /* :1803 */
{
    RX_GATE(p, ORC_PI / 2, a); RX_GATE(p, ORC_PI / 2, b); CX_GATE(p, a, b); RZ_GATE(p, theta, b);
    CX_GATE(p, a, b); RX_GATE(p, -ORC_PI / 2, a); RX_GATE(p, -ORC_PI / 2, b);
}

/* *_OP wrappers: which Gate field feeds which parameter, src/dmsim_nvgpu_omp.cuh:1821-2008 */
static int exe_op(pass_t* p, const orc_gate* g, const double* mats)
{
    const int q0 = g->qb[0], q1 = g->qb[1], q2 = g->qb[2], q3 = g->qb[3], q4 = g->qb[4];
    switch (g->op)
    {
    case OP_U3: U3_GATE(p, g->theta, g->phi, g->lambda, q0); break;
    case OP_U2: U2_GATE(p, g->phi, g->lambda, q0); break;
    case OP_U1: U1_GATE(p, g->lambda, q0); break;
    case OP_CX: CX_GATE(p, q0, q1); break;
    case OP_ID: break; /* ID_GATE :1272-1275 */
    case OP_X: X_GATE(p, q0); break;
    case OP_Y: Y_GATE(p, q0); break;
    case OP_Z: Z_GATE(p, q0); break;
    case OP_H: H_GATE(p, q0); break;
    case OP_S: S_GATE(p, q0); break;
    case OP_SDG: SDG_GATE(p, q0); break;
    case OP_T: T_GATE(p, q0); break;
    case OP_TDG: TDG_GATE(p, q0); break;
    case OP_RX: RX_GATE(p, g->theta, q0); break;
    case OP_RY: RY_GATE(p, g->theta, q0); break;
    case OP_RZ: RZ_GATE(p, g->phi, q0); break;
    case OP_CZ: CZ_GATE(p, q0, q1); break;
    case OP_CY: CY_GATE(p, q0, q1); break;
    case OP_SWAP: SWAP_GATE(p, q0, q1); break;
    case OP_CH: CH_GATE(p, q0, q1); break;
    case OP_CCX: CCX_GATE(p, q0, q1, q2); break;
    case OP_CSWAP: CSWAP_GATE(p, q0, q1, q2); break;
    case OP_CRX: CRX_GATE(p, g->lambda, q0, q1); break;
    case OP_CRY: CRY_GATE(p, g->lambda, q0, q1); break;
    case OP_CRZ: CRZ_GATE(p, g->lambda, q0, q1); break;
    case OP_CU1: CU1_GATE(p, g->lambda, q0, q1); break;
    case OP_CU3: CU3_GATE(p, g->theta, g->phi, g->lambda, q0, q1); break;
    case OP_RXX: RXX_GATE(p, g->theta, q0, q1); break;
    case OP_RZZ: RZZ_GATE(p, g->theta, q0, q1); break;
    case OP_RCCX: RCCX_GATE(p, q0, q1, q2); break;
    case OP_RC3X: RC3X_GATE(p, q0, q1, q2, q3); break;
    case OP_C3X: C3X_GATE(p, q0, q1, q2, q3); break;
    case OP_C3SQRTX: C3SQRTX_GATE(p, q0, q1, q2, q3); break;
    case OP_C4X: C4X_GATE(p, q0, q1, q2, q3, q4); break;
    case OP_R: R_GATE(p, g->theta, q0); break;
    case OP_SRN: SRN_GATE(p, q0); break;
    case OP_W: W_GATE(p, q0); break;
    case OP_RYY: RYY_GATE(p, g->theta, q0, q1); break;
    case OP_RAW_C1:
    {
        const double* m = mats + 32 * (size_t)g->mat;
        C1_GATE(p, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], q0);
        break;
    }
    case OP_RAW_C2: C2_GATE(p, mats + 32 * (size_t)g->mat, q0, q1); break;
    default: return -1;
    }
    return 0;
}

/* circuit(), src/dmsim_nvgpu_omp.cuh:816-822 */
static int circuit(pass_t* p, const orc_gate* gates, size_t n_gates, const double* mats)
{
    for (size_t t = 0; t < n_gates; t++)
        if (exe_op(p, &gates[t], mats)) return -1;
    return 0;
}

/* block_transpose with conjugation, GPU semantics, src/dmsim_nvgpu_omp.cuh:825-855 (1-GPU case):
 * out[x*dim + y] = conj(in[y*dim + x]) */
static void adjoint(const orc_sim_t* s, double* ore, double* oim, const double* ire, const double* iim)
{
    const idx_t dim = s->dim;
#pragma omp parallel for schedule(static)
    for (idx_t y = 0; y < dim; y++)
        for (idx_t x = 0; x < dim; x++)
        {
            ore[x * dim + y] = ire[y * dim + x];
            oim[x * dim + y] = -iim[y * dim + x];
        }
}

/* ---------------------------------- public C API ---------------------------------- */

void* orc_create(int n_qubits)
{
    orc_sim_t* s = (orc_sim_t*)calloc(1, sizeof(orc_sim_t));
    s->n = n_qubits;
    s->dim = (idx_t)1 << n_qubits;
    size_t bytes = (size_t)s->dim * s->dim * sizeof(double);
    s->re = (double*)calloc(1, bytes); s->im = (double*)calloc(1, bytes);
    s->bre = (double*)calloc(1, bytes); s->bim = (double*)calloc(1, bytes);
    if (!s->re || !s->im || !s->bre || !s->bim) return NULL;
    s->re[0] = 1.0; /* rho[0][0] = 1, src/dmsim_nvgpu_omp.cuh:239-240 */
    return s;
}

void orc_destroy(void* h)
{
    orc_sim_t* s = (orc_sim_t*)h;
    if (!s) return;
    free(s->re); free(s->im); free(s->bre); free(s->bim); free(s);
}

/* reset_dm(), src/dmsim_nvgpu_omp.cuh:308-329 */
void orc_reset(void* h)
{
    orc_sim_t* s = (orc_sim_t*)h;
    size_t bytes = (size_t)s->dim * s->dim * sizeof(double);
    memset(s->re, 0, bytes); memset(s->im, 0, bytes);
    memset(s->bre, 0, bytes); memset(s->bim, 0, bytes);
    s->re[0] = 1.0;
}

void orc_set_state(void* h, const double* re, const double* im)
{
    orc_sim_t* s = (orc_sim_t*)h;
    size_t bytes = (size_t)s->dim * s->dim * sizeof(double);
    memcpy(s->re, re, bytes); memcpy(s->im, im, bytes);
}

/* sim(), src/dmsim_nvgpu_omp.cuh:390-494 + simulation_kernel :918-939 (single-GPU branch):
 * forward circuit on dm; adjoint dm -> buf; circuit on buf; swap (buf becomes the state/result). */
int orc_sim(void* h, const orc_gate* gates, size_t n_gates, const double* mats)
{
    orc_sim_t* s = (orc_sim_t*)h;
    pass_t p = {s->n, s->dim, s->re, s->im};
    if (circuit(&p, gates, n_gates, mats)) return -1;
    adjoint(s, s->bre, s->bim, s->re, s->im);
    pass_t b = {s->n, s->dim, s->bre, s->bim};
    if (circuit(&b, gates, n_gates, mats)) return -1;
    double* t;
    t = s->re; s->re = s->bre; s->bre = t; /* swap_pointers :456-457 */
    t = s->im; s->im = s->bim; s->bim = t;
    return 0;
}

/* dm_real_res / dm_imag_res, src/dmsim_nvgpu_omp.cuh:458-459 */
void orc_get_dm(void* h, double* re, double* im)
{
    orc_sim_t* s = (orc_sim_t*)h;
    size_t bytes = (size_t)s->dim * s->dim * sizeof(double);
    if (re) memcpy(re, s->re, bytes);
    if (im) memcpy(im, s->im, bytes);
}

void orc_get_diag(void* h, double* diag)
{
    orc_sim_t* s = (orc_sim_t*)h;
    for (idx_t i = 0; i < s->dim; i++) diag[i] = s->re[i * s->dim + i];
}

/* measure(), src/dmsim_nvgpu_omp.cuh:521-549, with the seed exposed (reference: srand(time(0))).
 * Returns the final prefix sum (the reference warns when it is > 1e-3 from 1). */
double orc_measure(void* h, unsigned seed, unsigned repetition, uint64_t* res_state)
{
    orc_sim_t* s = (orc_sim_t*)h;
    const idx_t sv_num = s->dim;
    double* sv_diag_scan = (double*)malloc((sv_num + 1) * sizeof(double));
    sv_diag_scan[0] = 0;
    for (idx_t i = 1; i < sv_num + 1; i++)
        sv_diag_scan[i] = sv_diag_scan[i - 1] + fabs(s->re[(i - 1) * s->dim + (i - 1)]);
    srand(seed);
    memset(res_state, 0, repetition * sizeof(uint64_t));
    for (unsigned i = 0; i < repetition; i++)
    {
        double r = (double)rand() / (double)RAND_MAX;
        for (idx_t j = 0; j < sv_num; j++)
            if (sv_diag_scan[j] <= r && r < sv_diag_scan[j + 1]) res_state[i] = j;
    }
    double total = sv_diag_scan[sv_num];
    free(sv_diag_scan);
    return total;
}

/* Same sampling rule on caller-provided uniform numbers (lets tests compare the device sampler). */
void orc_sample_with_r(void* h, const double* r, unsigned repetition, uint64_t* res_state)
{
    orc_sim_t* s = (orc_sim_t*)h;
    const idx_t sv_num = s->dim;
    double* scan = (double*)malloc((sv_num + 1) * sizeof(double));
    scan[0] = 0;
    for (idx_t i = 1; i < sv_num + 1; i++) scan[i] = scan[i - 1] + fabs(s->re[(i - 1) * s->dim + (i - 1)]);
    memset(res_state, 0, repetition * sizeof(uint64_t));
    for (unsigned i = 0; i < repetition; i++)
        for (idx_t j = 0; j < sv_num; j++)
            if (scan[j] <= r[i] && r[i] < scan[j + 1]) res_state[i] = j;
    free(scan);
}
