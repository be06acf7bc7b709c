// include/DmSimApi.hpp -- the reference's XACC plugin ABI (xacc/DmSimApi.hpp:1-62) served by the B200 engine.
//
// XACC's DmSimAccelerator talks to a backend through four virtual calls (init / addGate / measure / finalize) and
// obtains it from the factory getGpuDmSim() (xacc/nvidia_omp/NvidiaOmpRunner.cu:8-177).  This header keeps
// namespace DmSim, the 38-value enum class OP (same order as DMSim::OP, xacc/DmSimApi.hpp:6-45) and the abstract
// class byte for byte in meaning, and implements the factory on include/dmsim_b200.hpp.  Header-only; link with
// -ldmsim_b200.  SURVEY.md section 8(f) rank 1.
//
// Differences from the reference runner, all deliberate:
//   * every OP of the enum is accepted (the reference hits __builtin_unreachable() for ID, CCX, CSWAP, CU3, RXX, RZZ,
//     RCCX, RC3X, C3X, C3SQRTX, C4X, R, SRN, W, RYY: NvidiaOmpRunner.cu:142-160);
//   * OP::CH appends CH (the reference appends SWAP by mistake, NvidiaOmpRunner.cu:106-110);
//   * wrong qubit / parameter counts are reported on stderr and abort (the reference asserts);
//   * measure() frees the shot array with delete[] (the reference uses scalar delete on a new[] array, :166).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#ifndef DMSIM_NO_PRINT_MEA
#define DMSIM_NO_PRINT_MEA // as the reference XACC runners: no per-sim() summary line on stdout
#endif
#include "dmsim_b200.hpp"

namespace DmSim
{
enum class OP
{
    U3, U2, U1, CX, ID, X, Y, Z, H, S,
    SDG, T, TDG, RX, RY, RZ, CZ, CY, SWAP, CH,
    CCX, CSWAP, CRX, CRY, CRZ, CU1, CU3, RXX, RZZ, RCCX,
    RC3X, C3X, C3SQRTX, C4X, R, SRN, W, RYY
};

class DmSimBackend
{
public:
    virtual void init(int n_qubits, int n_gpus = 1) = 0;
    virtual void addGate(OP op, const std::vector<int>& qubits, const std::vector<double>& params = {}) = 0;
    virtual std::vector<int64_t> measure(int shots) = 0;
    virtual void finalize() = 0;
    virtual ~DmSimBackend() {}
};

// The B200 engine behind the plugin ABI.  init(n_qubits, n_gpus) drives n_gpus devices of this process, like the
// reference runner (xacc/nvidia_omp/NvidiaOmpRunner.cu: one Simulation(n_qubits, n_gpus) behind the 4 calls).
class B200Backend : public DmSimBackend
{
public:
    void init(int n_qubits, int n_gpus = 1) override
    {
        m_sim = std::make_shared<::DMSim::Simulation>((::DMSim::IdxType)n_qubits, (::DMSim::IdxType)n_gpus);
    }
    void addGate(OP op, const std::vector<int>& qubits, const std::vector<double>& params = {}) override
    {
        // number of qubit / parameter operands of every OP, in enum order (reference src/dmsim_nvgpu_omp.cuh:580-767)
        static const unsigned char nq[38] = {1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2,
                                             3, 3, 2, 2, 2, 2, 2, 2, 2, 3, 4, 4, 4, 5, 1, 1, 1, 2};
        static const unsigned char np[38] = {3, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0,
                                             0, 0, 1, 1, 1, 1, 3, 1, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1};
        const int o = (int)op;
        if (!m_sim || o < 0 || o >= 38 || qubits.size() != nq[o] || params.size() != np[o])
        {
            fprintf(stderr, "Error: DmSimBackend::addGate(op=%d) with %zu qubits / %zu parameters%s\n", o, qubits.size(),
                    params.size(), m_sim ? "" : " before init()");
            exit(1);
        }
        typedef ::DMSim::Simulation S;
        const std::vector<int>& q = qubits;
        const std::vector<double>& p = params;
        ::DMSim::Gate* g = NULL;
        switch (op)
        {
        case OP::U3: g = S::U3(p[0], p[1], p[2], q[0]); break;
        case OP::U2: g = S::U2(p[0], p[1], q[0]); break;
        case OP::U1: g = S::U1(p[0], q[0]); break;
        case OP::CX: g = S::CX(q[0], q[1]); break;
        case OP::ID: g = S::ID(q[0]); break;
        case OP::X: g = S::X(q[0]); break;
        case OP::Y: g = S::Y(q[0]); break;
        case OP::Z: g = S::Z(q[0]); break;
        case OP::H: g = S::H(q[0]); break;
        case OP::S: g = S::S(q[0]); break;
        case OP::SDG: g = S::SDG(q[0]); break;
        case OP::T: g = S::T(q[0]); break;
        case OP::TDG: g = S::TDG(q[0]); break;
        case OP::RX: g = S::RX(p[0], q[0]); break;
        case OP::RY: g = S::RY(p[0], q[0]); break;
        case OP::RZ: g = S::RZ(p[0], q[0]); break;
        case OP::CZ: g = S::CZ(q[0], q[1]); break;
        case OP::CY: g = S::CY(q[0], q[1]); break;
        case OP::SWAP: g = S::SWAP(q[0], q[1]); break;
        case OP::CH: g = S::CH(q[0], q[1]); break;
        case OP::CCX: g = S::CCX(q[0], q[1], q[2]); break;
        case OP::CSWAP: g = S::CSWAP(q[0], q[1], q[2]); break;
        case OP::CRX: g = S::CRX(p[0], q[0], q[1]); break;
        case OP::CRY: g = S::CRY(p[0], q[0], q[1]); break;
        case OP::CRZ: g = S::CRZ(p[0], q[0], q[1]); break;
        case OP::CU1: g = S::CU1(p[0], q[0], q[1]); break;
        case OP::CU3: g = S::CU3(p[0], p[1], p[2], q[0], q[1]); break;
        case OP::RXX: g = S::RXX(p[0], q[0], q[1]); break;
        case OP::RZZ: g = S::RZZ(p[0], q[0], q[1]); break;
        case OP::RCCX: g = S::RCCX(q[0], q[1], q[2]); break;
        case OP::RC3X: g = S::RC3X(q[0], q[1], q[2], q[3]); break;
        case OP::C3X: g = S::C3X(q[0], q[1], q[2], q[3]); break;
        case OP::C3SQRTX: g = S::C3SQRTX(q[0], q[1], q[2], q[3]); break;
        case OP::C4X: g = S::C4X(q[0], q[1], q[2], q[3], q[4]); break;
        case OP::R: g = S::R(p[0], q[0]); break;
        case OP::SRN: g = S::SRN(q[0]); break;
        case OP::W: g = S::W(q[0]); break;
        case OP::RYY: g = S::RYY(p[0], q[0], q[1]); break;
        }
        m_sim->append(g); // deep copy (reference :331-344)
        delete g;
    }
    // upload + sim + measure, as the reference runner does (NvidiaOmpRunner.cu:153-170)
    std::vector<int64_t> measure(int shots) override
    {
        if (!m_sim)
        {
            fprintf(stderr, "Error: DmSimBackend::measure() before init()\n");
            exit(1);
        }
        m_sim->upload();
        m_sim->sim();
        ::DMSim::IdxType* res = m_sim->measure((unsigned)shots);
        std::vector<int64_t> result;
        result.reserve(shots);
        for (int i = 0; i < shots; ++i) result.emplace_back((int64_t)res[i]);
        delete[] res;
        return result;
    }
    void finalize() override { m_sim.reset(); }
    // non-breaking extra: the exact diagonal, for callers that want probabilities instead of shots
    std::vector<double> probabilities()
    {
        std::vector<double> d((size_t)1 << m_sim->n_qubits);
        m_sim->get_diag(d.data());
        return d;
    }

private:
    std::shared_ptr<::DMSim::Simulation> m_sim;
};

#ifndef XACC_HAS_CUDA
#define XACC_HAS_CUDA 1
#endif
inline std::shared_ptr<DmSimBackend> getGpuDmSim() { return std::make_shared<B200Backend>(); }
} // namespace DmSim
