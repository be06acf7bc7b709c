// include/dmsim_b200.hpp -- DM-Sim's C++ API (namespace DMSim: Gate, Simulation, print helpers) implemented on
// the B200-native C-ABI (include/dmsim_b200.h).  Header-only; link with -ldmsim_b200.
//
// Drop-in for the reference's GPU backend header (src/dmsim_nvgpu_omp.cuh together with src/config.hpp and
// src/util_nvgpu.cuh): same namespace, type aliases, enum OP order, Gate fields and dump() text, the 38 static gate
// factories with the same parameter order and parameter->field mapping (:580-767), append (deep copy, qubit-range
// asserts :331-344), upload / sim / clear_circuit / reset / reset_dm, measure (caller delete[]s, :521-549),
// dump(), print_res_sv / print_res_dm, the public result arrays dm_real_res / dm_imag_res (rho^T, [col][row]) and
// print_measurement / print_binary / cpu_timer / swap_pointers / is_power_of_2 from util_nvgpu.cuh.
//
// Differences, all deliberate:
//   * IdxType is 64-bit (the reference needs a patched config.hpp for n >= 15, src/config.hpp:43).
//   * The full-matrix D2H the reference does after every sim() (:458-466) is lazy: dm_real_res / dm_imag_res are
//     accessor calls res_real() / res_imag() that fetch on first use after a run (print_res_* and the public
//     pointers dm_real_res / dm_imag_res are refreshed by sync_results()).  measure() runs on the device.
//   * Simulation(n_qubits, n_gpus) drives all n_gpus devices from this one process like the reference (:196-271,
//     :397-399); the extra overload Simulation(n, world_size, rank, nccl_id) is the one-process-per-GPU form.
//   * Extra factories C1 / C2 expose the reference's unreachable generic gates (C1_GATE :1004, C2_GATE :1028).
//   * Errors: fatal like the reference (message on stderr + exit(1)); no exceptions cross the API.
#ifndef DMSIM_B200_HPP
#define DMSIM_B200_HPP

#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>

#include <complex>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "dmsim_b200.h"

// reference src/config.hpp:25: print the per-sim() summary line.  Define DMSIM_NO_PRINT_MEA before including this
// header to silence it (what the XACC runners do with `#undef PRINT_MEA_PER_CIRCUIT`,
// xacc/cpu_omp/dm_sim_omp_runner.cpp:2-4).
#if !defined(PRINT_MEA_PER_CIRCUIT) && !defined(DMSIM_NO_PRINT_MEA)
#define PRINT_MEA_PER_CIRCUIT
#endif

namespace DMSim
{
// ---- src/config.hpp ----
using IdxType = unsigned long long;
using ValType = double;
#ifndef RAND_SEED
#define RAND_SEED time(0)
#endif
#define DMSIM_ERROR_BAR (1e-3)
#ifndef PI
#define PI 3.14159265358979323846
#endif
#ifndef S2I
#define S2I 0.70710678118654752440
#endif

// ---- enum OP, src/dmsim_nvgpu_omp.cuh:42-48 (C1/C2 appended far after RYY, nothing renumbered) ----
enum OP
{
    U3, U2, U1, CX, ID, X, Y, Z, H, S,
    SDG, T, TDG, RX, RY, RZ, CZ, CY, SWAP, CH,
    CCX, CSWAP, CRX, CRY, CRZ, CU1, CU3, RXX, RZZ, RCCX,
    RC3X, C3X, C3SQRTX, C4X, R, SRN, W, RYY,
    C1 = 100, C2 = 101
};
static const char* const OP_NAMES[] = {"U3", "U2", "U1", "CX", "ID", "X", "Y", "Z", "H", "S",
                                       "SDG", "T", "TDG", "RX", "RY", "RZ", "CZ", "CY", "SWAP", "CH",
                                       "CCX", "CSWAP", "CRX", "CRY", "CRZ", "CU1", "CU3", "RXX", "RZZ", "RCCX",
                                       "RC3X", "C3X", "C3SQRTX", "C4X", "R", "SRN", "W", "RYY"};

// ---- src/util_nvgpu.cuh ----
inline void print_binary(IdxType v, int width)
{
    for (int i = width - 1; i >= 0; i--) putchar('0' + ((v >> i) & 1));
}
inline void print_measurement(IdxType* res_state, IdxType n_qubits, int repetition)
{
    assert(res_state != NULL);
    printf("\n===============  Measurement (tests=%d) ================\n", repetition);
    for (int i = 0; i < repetition; i++)
    {
        printf("Test-%d: ", i);
        print_binary(res_state[i], (int)n_qubits);
        printf("\n");
    }
}
inline void swap_pointers(ValType** pa, ValType** pb)
{
    ValType* tmp = (*pa);
    (*pa) = (*pb);
    (*pb) = tmp;
}
inline bool is_power_of_2(int x) { return (x > 0 && !(x & (x - 1))); }
inline double get_cpu_timer()
{
    struct timeval tp;
    gettimeofday(&tp, NULL);
    return ((double)tp.tv_sec + (double)tp.tv_usec * 1e-6) * 1e3; // ms
}
typedef struct CPU_TIMER
{
    CPU_TIMER() { start = stop = 0.0; }
    void start_timer() { start = get_cpu_timer(); }
    void stop_timer() { stop = get_cpu_timer(); }
    double measure() { return stop - start; }
    double start, stop;
} cpu_timer;

#define CHECK_NULL_POINTER(X) DMSim::__checkNullPointer(__FILE__, __LINE__, (void**)&(X))
inline void __checkNullPointer(const char* file, const int line, void** ptr)
{
    if ((*ptr) == NULL)
    {
        fprintf(stderr, "Error: NULL pointer at %s:%i.\n", file, line);
        exit(-1);
    }
}
#if defined(__CUDACC__) || defined(__CUDA_RUNTIME_H__)
// compiled by nvcc / after <cuda_runtime.h>, like the reference's users: the CUDA helpers of util_nvgpu.cuh:29-136 as they are there
#define cudaSafeCall(err) DMSim::__cudaSafeCall(err, __FILE__, __LINE__)
inline void __cudaSafeCall(cudaError_t err, const char* file, const int line)
{
    if (cudaSuccess != err)
    {
        fprintf(stderr, "cudaSafeCall() failed at %s:%i : %s\n", file, line, cudaGetErrorString(err));
        exit(-1);
    }
}
#define cudaCheckError() DMSim::__cudaCheckError(__FILE__, __LINE__)
inline void __cudaCheckError(const char* file, const int line)
{
    cudaError_t err = cudaGetLastError();
    if (cudaSuccess == err) err = cudaDeviceSynchronize();
    if (cudaSuccess != err)
    {
        fprintf(stderr, "cudaCheckError() failed at %s:%i : %s\n", file, line, cudaGetErrorString(err));
        exit(-1);
    }
}
#define SAFE_ALOC_HOST(X, Y) cudaSafeCall(cudaMallocHost((void**)&(X), (Y)));
#define SAFE_ALOC_GPU(X, Y) cudaSafeCall(cudaMalloc((void**)&(X), (Y)));
#define SAFE_FREE_HOST(X) if ((X) != NULL) { cudaSafeCall(cudaFreeHost((X))); (X) = NULL; }
#define SAFE_FREE_GPU(X) if ((X) != NULL) { cudaSafeCall(cudaFree((X))); (X) = NULL; }
typedef struct GPU_Timer
{
    GPU_Timer() { cudaSafeCall(cudaEventCreate(&this->start)); cudaSafeCall(cudaEventCreate(&this->stop)); }
    ~GPU_Timer() { cudaEventDestroy(this->start); cudaEventDestroy(this->stop); }
    void start_timer() { cudaSafeCall(cudaEventRecord(this->start)); }
    void stop_timer() { cudaSafeCall(cudaEventRecord(this->stop)); }
    double measure()
    {
        cudaSafeCall(cudaEventSynchronize(this->stop));
        float ms = 0;
        cudaSafeCall(cudaEventElapsedTime(&ms, this->start, this->stop));
        return (double)ms;
    }
    cudaEvent_t start, stop;
} gpu_timer;
#else
// host-only build (g++, no CUDA headers): gpu_timer keeps its name and meaning -- device work between start_timer() and
// stop_timer() -- by synchronising the device through the C ABI around a host clock
typedef struct GPU_Timer
{
    GPU_Timer() { start = stop = 0.0; }
    void start_timer() { dmb_device_synchronize(); start = get_cpu_timer(); }
    void stop_timer() { dmb_device_synchronize(); stop = get_cpu_timer(); }
    double measure() { return stop - start; }
    double start, stop;
} gpu_timer;
#endif

#define DMSIM_CHECK(call)                                                                     \
    do                                                                                        \
    {                                                                                         \
        int rc_ = (call);                                                                     \
        if (rc_ != DMB_OK)                                                                    \
        {                                                                                     \
            fprintf(stderr, "DM-Sim(b200) error %d at %s:%d: %s\n", rc_, __FILE__, __LINE__,  \
                    dmb_last_error());                                                        \
            exit(1);                                                                          \
        }                                                                                     \
    } while (0)

// ---- class Gate, src/dmsim_nvgpu_omp.cuh:99-191 (no device function pointer) ----
class Gate
{
public:
    Gate(enum OP _op_name, IdxType _qb0, IdxType _qb1, IdxType _qb2, IdxType _qb3, IdxType _qb4, ValType _theta,
         ValType _phi, ValType _lambda)
        : op_name(_op_name), qb0(_qb0), qb1(_qb1), qb2(_qb2), qb3(_qb3), qb4(_qb4), theta(_theta), phi(_phi),
          lambda(_lambda)
    {
    }
    ~Gate() {}
    void dump(std::stringstream& ss)
    {
        const char* nm = op_name == OP::C1 ? "C1" : (op_name == OP::C2 ? "C2" : OP_NAMES[op_name]);
        ss << nm << "(" << qb0 << "," << qb1 << "," << qb2 << "," << qb3 << "," << qb4 << "," << theta << "," << phi
           << "," << lambda << ");" << std::endl;
    }
    enum OP op_name;
    IdxType qb0, qb1, qb2, qb3, qb4;
    ValType theta, phi, lambda;
    std::vector<std::complex<double>> matrix; // C1 (4 entries) / C2 (16 entries, index 2*bit(qb0)+bit(qb1)) only
};

// ---- class Simulation, src/dmsim_nvgpu_omp.cuh:193-814 ----
class Simulation
{
public:
    // n_gpus > 1: ONE process, devices 0 .. n_gpus-1 with peer access between every pair (reference :259-269)
    Simulation(IdxType _n_qubits, IdxType _n_gpus) { init(_n_qubits, _n_gpus, _n_gpus > 1 ? (long long)DMB_ALL_RANKS : 0, NULL); }
    // one process per GPU: rank r of world size _n_gpus; nccl_id = the 128 bytes from dmb_comm_unique_id on rank 0
    Simulation(IdxType _n_qubits, IdxType _n_gpus, IdxType rank, const uint8_t* nccl_id) { init(_n_qubits, _n_gpus, (long long)rank, nccl_id); }
    ~Simulation()
    {
        clear_circuit();
        dmb_destroy(h);
        free(dm_real_res);
        free(dm_imag_res);
    }
    Simulation(const Simulation&) = delete;
    Simulation& operator=(const Simulation&) = delete;

    void reset()
    {
        clear_circuit();
        reset_dm();
    }
    void reset_dm()
    {
        DMSIM_CHECK(dmb_reset_dm(h));
        results_valid = false;
    }
    // add a gate to the current circuit (deep copy: the caller keeps ownership, reference :331-344)
    void append(Gate* g)
    {
        if (g == NULL)
        {
            fprintf(stderr, "Error: pointer g is null! (dmsim_b200)\n");
            exit(-1);
        }
        assert((g->qb0 < n_qubits));
        assert((g->qb1 < n_qubits));
        assert((g->qb2 < n_qubits));
        assert((g->qb3 < n_qubits));
        assert((g->qb4 < n_qubits));
        circuit.push_back(new Gate(*g));
        n_gates++;
    }
    Simulation* upload()
    {
        assert(n_gates == circuit.size());
        assert(!uploaded); // reference asserts circuit_gpu == NULL (:349-350): clear_circuit() first
        std::vector<dmb_gate> rec(n_gates);
        std::vector<double> mats;
        for (IdxType t = 0; t < n_gates; t++)
        {
            const Gate& g = *circuit[t];
            dmb_gate& r = rec[t];
            memset(&r, 0, sizeof(r));
            r.op = (int32_t)g.op_name;
            r.qb[0] = (int32_t)g.qb0; r.qb[1] = (int32_t)g.qb1; r.qb[2] = (int32_t)g.qb2;
            r.qb[3] = (int32_t)g.qb3; r.qb[4] = (int32_t)g.qb4;
            r.theta = g.theta; r.phi = g.phi; r.lambda = g.lambda;
            if (g.op_name == OP::C1 || g.op_name == OP::C2)
            {
                r.mat = (int64_t)(mats.size() / 32);
                mats.resize(mats.size() + 32, 0.0);
                double* slot = mats.data() + mats.size() - 32;
                for (size_t e = 0; e < g.matrix.size() && e < 16; e++)
                {
                    slot[2 * e] = g.matrix[e].real();
                    slot[2 * e + 1] = g.matrix[e].imag();
                }
            }
        }
        DMSIM_CHECK(dmb_set_circuit(h, rec.data(), rec.size(), mats.empty() ? NULL : mats.data(), mats.size() / 32));
        uploaded = true;
        return this;
    }
    std::string dump()
    {
        std::stringstream ss;
        for (IdxType t = 0; t < n_gates; t++) circuit[t]->dump(ss);
        return ss.str();
    }
    // start dm simulation (blocking), reference :390-494
    void sim()
    {
        dmb_stats st;
        DMSIM_CHECK(dmb_run(h, &st));
        results_valid = false;
        last_stats = st;
#ifdef PRINT_MEA_PER_CIRCUIT
        // same fields as the reference's summary line (:484-490)
        const double mem_mb = (double)dm_size / 1024.0 / 1024.0 * (double)(2 + (st.n_exchanges ? 2 : 0));
        printf("\n============== DM-Sim ===============\n");
        printf("nqubits:%d, ngates:%d, ngpus:%d, comp:%.3lf ms, comm:%.3lf ms, sim:%.3lf ms, mem:%.3lf MB, mem_per_gpu:%.3lf MB\n",
               (int)n_qubits, (int)n_gates, (int)n_gpus, st.comp_ms, st.comm_ms, st.sim_ms, mem_mb, mem_mb / (double)n_gpus);
        printf("=====================================\n");
#endif
    }
    void clear_circuit()
    {
        for (IdxType i = 0; i < circuit.size(); i++) delete circuit[i];
        circuit.clear();
        n_gates = 0;
        uploaded = false;
        if (h) dmb_clear_circuit(h);
    }
    // reference :521-549; caller delete[]s the result
    IdxType* measure(unsigned repetition = 10)
    {
        IdxType* res_state = new IdxType[repetition];
        std::vector<uint64_t> out(repetition, 0);
        double total = 0.0;
        DMSIM_CHECK(dmb_measure(h, (unsigned)(RAND_SEED), repetition, out.data(), &total));
        for (unsigned i = 0; i < repetition; i++) res_state[i] = (IdxType)out[i];
        double diff = total - 1.0;
        if (diff < 0) diff = -diff;
        if (diff > DMSIM_ERROR_BAR) printf("Sum of probability along diag is far from 1.0 with %lf\n", total);
        return res_state;
    }
    // fetch dm_real_res / dm_imag_res (the reference copies them after every sim(); here on demand)
    void sync_results()
    {
        if (results_valid) return;
        if (!dm_real_res)
        {
            dm_real_res = (ValType*)malloc(dm_size);
            dm_imag_res = (ValType*)malloc(dm_size);
            if (!dm_real_res || !dm_imag_res)
            {
                fprintf(stderr, "Error: host allocation of the result arrays failed\n");
                exit(1);
            }
        }
        DMSIM_CHECK(dmb_get_dm(h, dm_real_res, dm_imag_res));
        results_valid = true;
    }
    const ValType* res_real() { sync_results(); return dm_real_res; }
    const ValType* res_imag() { sync_results(); return dm_imag_res; }
    void print_res_sv()
    {
        std::vector<double> d(dim);
        DMSIM_CHECK(dmb_get_diag(h, d.data()));
        printf("----- Real SV ------\n");
        for (IdxType i = 0; i < dim; i++) printf("%lf ", d[i]);
        printf("\n");
        sync_results();
        printf("----- Imag SV ------\n");
        for (IdxType i = 0; i < dim; i++) printf("%lf ", dm_imag_res[i * dim + i]);
        printf("\n");
    }
    void print_res_dm()
    {
        sync_results();
        printf("----- Real DM------\n");
        for (IdxType i = 0; i < dim; i++)
        {
            for (IdxType j = 0; j < dim; j++) printf("%lf ", dm_real_res[i * dim + j]);
            printf("\n");
        }
        printf("----- Imag DM------\n");
        for (IdxType i = 0; i < dim; i++)
        {
            for (IdxType j = 0; j < dim; j++) printf("%lf ", dm_imag_res[i * dim + j]);
            printf("\n");
        }
    }
    // non-breaking extras
    double trace() { double v = 0; DMSIM_CHECK(dmb_trace(h, &v)); return v; }
    double purity() { double v = 0; DMSIM_CHECK(dmb_purity(h, &v)); return v; }
    void get_diag(double* out) { DMSIM_CHECK(dmb_get_diag(h, out)); }
    dmb_handle handle() { return h; }

    // =============================== Standard Gates (reference :580-767) ===================================
    static Gate* U3(ValType theta, ValType phi, ValType lambda, IdxType m) { return new Gate(OP::U3, m, 0, 0, 0, 0, theta, phi, lambda); }
    static Gate* U2(ValType phi, ValType lambda, IdxType m) { return new Gate(OP::U2, m, 0, 0, 0, 0, 0., phi, lambda); }
    static Gate* U1(ValType lambda, IdxType m) { return new Gate(OP::U1, m, 0, 0, 0, 0, 0., 0., lambda); }
    static Gate* CX(IdxType m, IdxType n) { return new Gate(OP::CX, m, n, 0, 0, 0, 0., 0., 0.); }
    static Gate* ID(IdxType m) { return new Gate(OP::ID, m, 0, 0, 0, 0, 0., 0., 0.); }
    static Gate* X(IdxType m) { return new Gate(OP::X, m, 0, 0, 0, 0, 0., 0., 0.); }
    static Gate* Y(IdxType m) { return new Gate(OP::Y, m, 0, 0, 0, 0, 0., 0., 0.); }
    static Gate* Z(IdxType m) { return new Gate(OP::Z, m, 0, 0, 0, 0, 0., 0., 0.); }
    static Gate* H(IdxType m) { return new Gate(OP::H, m, 0, 0, 0, 0, 0., 0., 0.); }
    static Gate* S(IdxType m) { return new Gate(OP::S, m, 0, 0, 0, 0, 0., 0., 0.); }
    static Gate* SDG(IdxType m) { return new Gate(OP::SDG, m, 0, 0, 0, 0, 0., 0., 0.); }
    static Gate* T(IdxType m) { return new Gate(OP::T, m, 0, 0, 0, 0, 0., 0., 0.); }
    static Gate* TDG(IdxType m) { return new Gate(OP::TDG, m, 0, 0, 0, 0, 0., 0., 0.); }
    static Gate* RX(ValType theta, IdxType m) { return new Gate(OP::RX, m, 0, 0, 0, 0, theta, 0., 0.); }
    static Gate* RY(ValType theta, IdxType m) { return new Gate(OP::RY, m, 0, 0, 0, 0, theta, 0., 0.); }
    static Gate* RZ(ValType phi, IdxType m) { return new Gate(OP::RZ, m, 0, 0, 0, 0, 0., phi, 0.); }
    static Gate* CZ(IdxType m, IdxType n) { return new Gate(OP::CZ, m, n, 0, 0, 0, 0., 0., 0.); }
    static Gate* CY(IdxType m, IdxType n) { return new Gate(OP::CY, m, n, 0, 0, 0, 0., 0., 0.); }
    static Gate* SWAP(IdxType m, IdxType n) { return new Gate(OP::SWAP, m, n, 0, 0, 0, 0., 0., 0.); }
    static Gate* CH(IdxType m, IdxType n) { return new Gate(OP::CH, m, n, 0, 0, 0, 0., 0., 0.); }
    static Gate* CCX(IdxType l, IdxType m, IdxType n) { return new Gate(OP::CCX, l, m, n, 0, 0, 0., 0., 0.); }
    static Gate* CSWAP(IdxType l, IdxType m, IdxType n) { return new Gate(OP::CSWAP, l, m, n, 0, 0, 0., 0., 0.); }
    static Gate* CRX(ValType lambda, IdxType m, IdxType n) { return new Gate(OP::CRX, m, n, 0, 0, 0, 0., 0., lambda); }
    static Gate* CRY(ValType lambda, IdxType m, IdxType n) { return new Gate(OP::CRY, m, n, 0, 0, 0, 0., 0., lambda); }
    static Gate* CRZ(ValType lambda, IdxType m, IdxType n) { return new Gate(OP::CRZ, m, n, 0, 0, 0, 0., 0., lambda); }
    static Gate* CU1(ValType lambda, IdxType m, IdxType n) { return new Gate(OP::CU1, m, n, 0, 0, 0, 0., 0., lambda); }
    static Gate* CU3(ValType theta, ValType phi, ValType lambda, IdxType m, IdxType n) { return new Gate(OP::CU3, m, n, 0, 0, 0, theta, phi, lambda); }
    static Gate* RXX(ValType theta, IdxType m, IdxType n) { return new Gate(OP::RXX, m, n, 0, 0, 0, theta, 0., 0.); }
    static Gate* RZZ(ValType theta, IdxType m, IdxType n) { return new Gate(OP::RZZ, m, n, 0, 0, 0, theta, 0., 0.); }
    static Gate* RCCX(IdxType l, IdxType m, IdxType n) { return new Gate(OP::RCCX, l, m, n, 0, 0, 0., 0., 0.); }
    static Gate* RC3X(IdxType l, IdxType m, IdxType n, IdxType o) { return new Gate(OP::RC3X, l, m, n, o, 0, 0., 0., 0.); }
    static Gate* C3X(IdxType l, IdxType m, IdxType n, IdxType o) { return new Gate(OP::C3X, l, m, n, o, 0, 0., 0., 0.); }
    static Gate* C3SQRTX(IdxType l, IdxType m, IdxType n, IdxType o) { return new Gate(OP::C3SQRTX, l, m, n, o, 0, 0., 0., 0.); }
    static Gate* C4X(IdxType l, IdxType m, IdxType n, IdxType o, IdxType p) { return new Gate(OP::C4X, l, m, n, o, p, 0., 0., 0.); }
    static Gate* R(ValType theta, IdxType m) { return new Gate(OP::R, m, 0, 0, 0, 0, theta, 0., 0.); }
    static Gate* SRN(IdxType m) { return new Gate(OP::SRN, m, 0, 0, 0, 0, 0., 0., 0.); }
    static Gate* W(IdxType m) { return new Gate(OP::W, m, 0, 0, 0, 0, 0., 0., 0.); }
    static Gate* RYY(ValType theta, IdxType m, IdxType n) { return new Gate(OP::RYY, m, n, 0, 0, 0, theta, 0., 0.); }
    // the reference's generic gates, made reachable: e = row-major 2x2 / 4x4 complex
    static Gate* C1(const std::complex<double>* e, IdxType m)
    {
        Gate* g = new Gate(OP::C1, m, 0, 0, 0, 0, 0., 0., 0.);
        g->matrix.assign(e, e + 4);
        return g;
    }
    static Gate* C2(const std::complex<double>* e, IdxType m, IdxType n)
    {
        Gate* g = new Gate(OP::C2, m, n, 0, 0, 0, 0., 0., 0.);
        g->matrix.assign(e, e + 16);
        return g;
    }

public:
    // n_qubits is the number of qubits
    IdxType n_qubits = 0;
    IdxType n_gpus = 1;
    IdxType dim = 0, half_dim = 0;
    IdxType dm_num = 0;  // 4^n
    IdxType dm_size = 0; // 8 * 4^n bytes per split array
    IdxType n_gates = 0;
    // results (rho^T, [col][row]); valid after sync_results()
    ValType* dm_real_res = NULL;
    ValType* dm_imag_res = NULL;
    std::vector<Gate*> circuit;
    dmb_stats last_stats;

private:
    void init(IdxType _n_qubits, IdxType _n_gpus, long long rank, const uint8_t* nccl_id)
    {
        n_qubits = _n_qubits;
        n_gpus = _n_gpus;
        dim = (IdxType)1 << n_qubits;
        half_dim = (IdxType)1 << (n_qubits - 1);
        dm_num = dim * dim;
        dm_size = dm_num * (IdxType)sizeof(ValType);
        memset(&last_stats, 0, sizeof(last_stats));
        // reference ctor (:218-229): power of two and dividing 2^n, else message + exit(1)
        if (!is_power_of_2((int)n_gpus))
        {
            std::cerr << "Error: Number of GPUs should be an exponential of 2." << std::endl;
            exit(1);
        }
        if (dim % n_gpus != 0)
        {
            std::cerr << "Error: Number of GPUs is too large or too small." << std::endl;
            exit(1);
        }
        DMSIM_CHECK(dmb_create((int)n_qubits, (int)n_gpus, (int)rank, -1, &h));
        if (n_gpus > 1 && rank != (long long)DMB_ALL_RANKS)
        {
            if (!nccl_id)
            {
                std::cerr << "Error: the one-process-per-GPU constructor needs the NCCL unique id of rank 0." << std::endl;
                exit(1);
            }
            DMSIM_CHECK(dmb_comm_init(h, nccl_id));
        }
    }
    dmb_handle h = NULL;
    bool uploaded = false;
    bool results_valid = false;
};

} // namespace DMSim
#endif // DMSIM_B200_HPP
