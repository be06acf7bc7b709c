// examples/xacc_backend.cpp -- the reference's XACC plugin ABI (xacc/DmSimApi.hpp) driven the way DmSimAccelerator
// does (xacc/DmSimAccelerator.cpp: init -> addGate* -> measure -> finalize), on the B200 backend.  Checks mirror
// xacc/nvidia_omp/tests/DmSimAcceleratorTester.cpp: Bell 0.5/0.5 (:8-27), <Z> = 1 - 2 sin^2(theta/2) for an RX
// sweep (:69-95) and the CH gate (which the reference runner maps to SWAP by mistake).
//     g++ -O2 -std=c++17 -I include examples/xacc_backend.cpp -L dm-sim_b200/lib -ldmsim_b200 -Wl,-rpath,dm-sim_b200/lib
#include <math.h>
#include <stdio.h>

#include "DmSimApi.hpp"

int main()
{
    const int shots = 8192;
    int bad = 0;
    auto backend = DmSim::getGpuDmSim();
    // Bell pair
    backend->init(2, 1);
    backend->addGate(DmSim::OP::H, {0});
    backend->addGate(DmSim::OP::CX, {0, 1});
    int n00 = 0, n11 = 0;
    for (int64_t s : backend->measure(shots)) { n00 += s == 0; n11 += s == 3; }
    backend->finalize();
    printf("bell: 00 %.4f 11 %.4f\n", (double)n00 / shots, (double)n11 / shots);
    bad += n00 + n11 != shots || fabs((double)n00 / shots - 0.5) > 0.05;
    // RX sweep
    for (int i = 0; i <= 4; i++)
    {
        const double theta = -M_PI + i * M_PI / 2.0;
        backend->init(1, 1);
        backend->addGate(DmSim::OP::RX, {0}, {theta});
        int n1 = 0;
        for (int64_t s : backend->measure(shots)) n1 += s == 1;
        backend->finalize();
        const double z = 1.0 - 2.0 * n1 / shots, want = 1.0 - 2.0 * sin(theta / 2) * sin(theta / 2);
        printf("rx(%+.4f): <Z> %+.4f expected %+.4f\n", theta, z, want);
        bad += fabs(z - want) > 0.05;
    }
    // CH: control |1> puts the target into |+>
    backend->init(2, 1);
    backend->addGate(DmSim::OP::X, {0});
    backend->addGate(DmSim::OP::CH, {0, 1});
    int n01 = 0; n11 = 0;
    for (int64_t s : backend->measure(shots)) { n01 += s == 1; n11 += s == 3; }
    backend->finalize();
    printf("ch: 01 %.4f 11 %.4f\n", (double)n01 / shots, (double)n11 / shots);
    bad += n01 + n11 != shots || fabs((double)n01 / shots - 0.5) > 0.05;
    // a gate the reference runner rejects
    backend->init(3, 1);
    backend->addGate(DmSim::OP::X, {0});
    backend->addGate(DmSim::OP::X, {1});
    backend->addGate(DmSim::OP::CCX, {0, 1, 2});
    int n7 = 0;
    for (int64_t s : backend->measure(64)) n7 += s == 7;
    backend->finalize();
    printf("ccx: 111 %d/64\n", n7);
    bad += n7 != 64;
    printf(bad ? "FAILED\n" : "OK\n");
    return bad != 0;
}
