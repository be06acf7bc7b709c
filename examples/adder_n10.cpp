// examples/adder_n10.cpp -- the reference's example/adder_n10_nvgpu_omp.cu driver (10-qubit Cuccaro adder, result
// 1000000010, README.md:228-246) written against the drop-in header.  Only the #include differs from the
// reference driver:
//     g++ -O2 -std=c++17 -I include examples/adder_n10.cpp -L dm-sim_b200/lib -ldmsim_b200 -Wl,-rpath,dm-sim_b200/lib
#include <stdio.h>

#include "dmsim_b200.hpp"

using namespace DMSim;

// majority / un-majority blocks of the ripple-carry adder
static void majority(Simulation& sim, const IdxType a, const IdxType b, const IdxType c)
{
    Gate* g;
    g = Simulation::CX(c, b); sim.append(g); delete g;
    g = Simulation::CX(c, a); sim.append(g); delete g;
    g = Simulation::CCX(a, b, c); sim.append(g); delete g;
}
static void unmaj(Simulation& sim, const IdxType a, const IdxType b, const IdxType c)
{
    Gate* g;
    g = Simulation::CCX(a, b, c); sim.append(g); delete g;
    g = Simulation::CX(c, a); sim.append(g); delete g;
    g = Simulation::CX(a, b); sim.append(g); delete g;
}

int main(int argc, char** argv)
{
    const int n_qubits = 10;
    // the reference example hard-codes n_gpus = 4 (example/adder_n10_nvgpu_omp.cu:45); here: ./adder_n10 [n_gpus], default 1
    const int n_gpus = argc > 1 ? atoi(argv[1]) : 1;
    const IdxType cin = 0, a[4] = {1, 2, 3, 4}, b[4] = {5, 6, 7, 8}, cout = 9;
    srand(time(0));
    Simulation sim(n_qubits, n_gpus);
    Gate* g;
    // a = 0001, b = 1111
    g = Simulation::X(a[0]); sim.append(g); delete g;
    for (int i = 0; i < 4; i++) { g = Simulation::X(b[i]); sim.append(g); delete g; }
    majority(sim, cin, b[0], a[0]);
    for (int i = 0; i < 3; i++) majority(sim, a[i], b[i + 1], a[i + 1]);
    g = Simulation::CX(a[3], cout); sim.append(g); delete g;
    for (int i = 2; i >= 0; i--) unmaj(sim, a[i], b[i + 1], a[i + 1]);
    unmaj(sim, cin, b[0], a[0]);
    sim.upload();
    sim.sim();
    IdxType* res = sim.measure(5);
    print_measurement(res, n_qubits, 5);
    int ok = 1;
    for (int i = 0; i < 5; i++) ok &= (res[i] == 0x202); // 1000000010
    delete[] res;
    printf("trace = %.15f, purity = %.15f -> %s\n", sim.trace(), sim.purity(), ok ? "OK" : "MISMATCH");
    return ok ? 0 : 1;
}
