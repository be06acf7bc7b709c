// dm-sim_b200/csrc/pybind_module.cpp -- pybind11 module with the reference's Python surface
// (src/py_nvgpu_omp_wrapper.cu:29-87): module name libdmsim_py_nvgpu_omp, classes Gate and Simulation with
// append / upload / clear_circuit / run / reset / measure(repetition) -> list and the 38 static factories.
// Scripts written for the reference (tool/dmsim_qasm.py output, example/adder_n10_omp.py) import it unchanged;
// the alias module dmsim_py_omp_wrapper re-exports it.  Extras: C1/C2 factories, get_dm, diag, trace, purity, dump.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#undef PRINT_MEA_PER_CIRCUIT
#define PRINT_MEA_PER_CIRCUIT
#include "dmsim_b200.hpp"

namespace py = pybind11;
using namespace DMSim;

PYBIND11_MODULE(libdmsim_py_nvgpu_omp, m)
{
    m.doc() = "DM-Sim Python API on the B200-native engine (drop-in for libdmsim_py_nvgpu_omp)";
    py::enum_<OP>(m, "OP")
        .value("U3", OP::U3).value("U2", OP::U2).value("U1", OP::U1).value("CX", OP::CX).value("ID", OP::ID)
        .value("X", OP::X).value("Y", OP::Y).value("Z", OP::Z).value("H", OP::H).value("S", OP::S)
        .value("SDG", OP::SDG).value("T", OP::T).value("TDG", OP::TDG).value("RX", OP::RX).value("RY", OP::RY)
        .value("RZ", OP::RZ).value("CZ", OP::CZ).value("CY", OP::CY).value("SWAP", OP::SWAP).value("CH", OP::CH)
        .value("CCX", OP::CCX).value("CSWAP", OP::CSWAP).value("CRX", OP::CRX).value("CRY", OP::CRY)
        .value("CRZ", OP::CRZ).value("CU1", OP::CU1).value("CU3", OP::CU3).value("RXX", OP::RXX)
        .value("RZZ", OP::RZZ).value("RCCX", OP::RCCX).value("RC3X", OP::RC3X).value("C3X", OP::C3X)
        .value("C3SQRTX", OP::C3SQRTX).value("C4X", OP::C4X).value("R", OP::R).value("SRN", OP::SRN)
        .value("W", OP::W).value("RYY", OP::RYY).value("C1", OP::C1).value("C2", OP::C2);

    py::class_<Gate>(m, "Gate")
        .def(py::init<enum OP, IdxType, IdxType, IdxType, IdxType, IdxType, ValType, ValType, ValType>())
        .def("dump", [](Gate& g) { std::stringstream ss; g.dump(ss); return ss.str(); });

    py::class_<Simulation>(m, "Simulation")
        .def(py::init<IdxType, IdxType>())
        .def("append", &Simulation::append)
        .def("upload", &Simulation::upload, py::return_value_policy::reference)
        .def("clear_circuit", &Simulation::clear_circuit)
        .def("run", &Simulation::sim)
        .def("reset", &Simulation::reset)
        .def("measure", [](Simulation& s, unsigned repetition) -> py::list {
            IdxType* m_rtn = s.measure(repetition);
            py::list rtn;
            for (unsigned i = 0; i < repetition; i++) rtn.append(m_rtn[i]);
            delete[] m_rtn;
            return rtn;
        })
        // extras (not in the reference)
        .def("dump", &Simulation::dump)
        .def("trace", &Simulation::trace)
        .def("purity", &Simulation::purity)
        .def("diag", [](Simulation& s) {
            py::array_t<double> d((py::ssize_t)s.dim);
            s.get_diag(d.mutable_data());
            return d;
        })
        .def("get_dm", [](Simulation& s) {
            s.sync_results();
            py::array_t<double> re({(py::ssize_t)s.dim, (py::ssize_t)s.dim}), im({(py::ssize_t)s.dim, (py::ssize_t)s.dim});
            memcpy(re.mutable_data(), s.dm_real_res, s.dm_size);
            memcpy(im.mutable_data(), s.dm_imag_res, s.dm_size);
            return py::make_tuple(re, im);
        })
        .def_static("U3", &Simulation::U3).def_static("U2", &Simulation::U2).def_static("U1", &Simulation::U1)
        .def_static("CX", &Simulation::CX).def_static("ID", &Simulation::ID).def_static("X", &Simulation::X)
        .def_static("Y", &Simulation::Y).def_static("Z", &Simulation::Z).def_static("H", &Simulation::H)
        .def_static("S", &Simulation::S).def_static("SDG", &Simulation::SDG).def_static("T", &Simulation::T)
        .def_static("TDG", &Simulation::TDG).def_static("RX", &Simulation::RX).def_static("RY", &Simulation::RY)
        .def_static("RZ", &Simulation::RZ).def_static("CZ", &Simulation::CZ).def_static("CY", &Simulation::CY)
        .def_static("SWAP", &Simulation::SWAP).def_static("CH", &Simulation::CH).def_static("CCX", &Simulation::CCX)
        .def_static("CSWAP", &Simulation::CSWAP).def_static("CRX", &Simulation::CRX).def_static("CRY", &Simulation::CRY)
        .def_static("CRZ", &Simulation::CRZ).def_static("CU1", &Simulation::CU1).def_static("CU3", &Simulation::CU3)
        .def_static("RXX", &Simulation::RXX).def_static("RZZ", &Simulation::RZZ).def_static("RCCX", &Simulation::RCCX)
        .def_static("RC3X", &Simulation::RC3X).def_static("C3X", &Simulation::C3X)
        .def_static("C3SQRTX", &Simulation::C3SQRTX).def_static("C4X", &Simulation::C4X).def_static("R", &Simulation::R)
        .def_static("SRN", &Simulation::SRN).def_static("W", &Simulation::W).def_static("RYY", &Simulation::RYY)
        .def_static("C1", [](std::vector<std::complex<double>> e, IdxType q) {
            if (e.size() != 4) throw std::invalid_argument("C1 needs 4 matrix entries (row-major 2x2)");
            return Simulation::C1(e.data(), q);
        })
        .def_static("C2", [](std::vector<std::complex<double>> e, IdxType q1, IdxType q2) {
            if (e.size() != 16) throw std::invalid_argument("C2 needs 16 matrix entries (row-major 4x4)");
            return Simulation::C2(e.data(), q1, q2);
        });
}
