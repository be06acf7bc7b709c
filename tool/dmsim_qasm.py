#!/usr/bin/env python3
"""dmsim_qasm.py -- DM_Sim Assembler for OpenQASM-V2.0: translates OpenQASM to a DM-Sim circuit script.

Python-3 replacement of the reference's tool/dmsim_qasm.py (which only runs under Python 2): same command line
(-i/--input, -o/--output, default dmsim_circuit.py), same generated script (it imports dmsim_py_omp_wrapper and is run
as `python circuit.py n_qubits n_gpus`) and the same statistics on stdout.  Extra: --run executes the circuit directly
on the engine without the script round trip.
"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main(argv=None):
    parser = argparse.ArgumentParser(description="DM_Sim Assembler for OpenQASM-V2.0: translating OpenQASM to DM_sim "
                                                 "native simulation circuit code.")
    parser.add_argument("--input", "-i", required=True, help="input OpenQASM file, such as adder.qasm")
    parser.add_argument("--output", "-o", default="dmsim_circuit.py",
                        help="output DM_Sim circuit python file (default: dmsim_circuit.py)")
    parser.add_argument("--run", action="store_true", help="also run the circuit on the GPU engine and print 10 shots")
    args = parser.parse_args(argv)
    qasm = importlib.import_module("dm-sim_b200.qasm")
    with open(args.input) as f:
        text = f.read()
    script, stats = qasm.translate(text)
    with open(args.output, "w") as f:
        f.write(script)
    print("== DM-Sim: Translating " + args.input + " to " + args.output + " ==")
    print("Number of qubits: " + str(stats["n_qubits"]))
    print("Number of basic gates: " + str(stats["basic_gates"]))
    print("Number of cnot gates: " + str(stats["cnot_gates"]))
    if args.run:
        dm = importlib.import_module("dm-sim_b200")
        n, gates = qasm.load(text)
        sim = dm.Simulation(n, 1)
        for g in gates:
            sim.append(dm.Gate(g[0], *(g[1] + [0] * (5 - len(g[1]))), theta=g[2], phi=g[3], lam=g[4]))
        sim.upload()
        sim.run()
        dm.print_measurement(sim.measure(10), n, 10)
    return 0


if __name__ == "__main__":
    sys.exit(main())
