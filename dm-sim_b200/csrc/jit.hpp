// dm-sim_b200/csrc/jit.hpp -- run-time specialised sweep kernels (see jit.cu).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "devop.hpp"

namespace dmb
{
struct JitKernel
{
    std::string key;           // hex hash of the generated text (also the file name in the disk cache)
    std::atomic<int> state{0}; // 0 compiling, 1 ready (cubin present), -1 failed (log says why)
    std::vector<char> cubin;
    std::string log;
    double compile_ms = 0;
    std::mutex m;              // guards the load below
    void* lib = nullptr;       // cudaLibrary_t
    void* kern = nullptr;      // cudaKernel_t
    unsigned long long attr_mask = 0; // devices on which the dynamic shared-memory limit has been raised
};

// the sweep's program as CUDA text: `defines` (structural literals) + `program` (the body of "dmb_jit_program.inc").
// `stream` / `rounds` / `groups` are the HOST copies of the sweep's device tables, `a` its filled parameter block.
// false: something the generator does not cover (the interpreter kernel runs the sweep)
bool jit_generate(const SweepArgs& a, const unsigned char* stream, const DevRound* rounds, const DevGroup* groups, std::string& defines,
                  std::string& program);
// hash of everything the generated text depends on (structural parameters, round / group tables, the ops' vid and aux) without
// generating it: the process-wide map  key -> cache entry  spares a re-planned circuit of known structure the text generation
bool jit_struct_key(const SweepArgs& a, const unsigned char* stream, const DevRound* rounds, const DevGroup* groups, unsigned long long (&key)[2]);
JitKernel* jit_lookup_struct(const unsigned long long (&key)[2], bool* known); // (a known structure may map to nullptr: not covered)
void jit_remember_struct(const unsigned long long (&key)[2], JitKernel* k);
bool jit_available(std::string* why);
// cache lookup by the hash of the text; a miss queues the compilation on the worker threads and returns the pending entry
JitKernel* jit_request(const std::string& defines, const std::string& program);
void jit_wait(JitKernel* k); // until the entry is ready or has failed
// launchable function of a ready entry on `device` (current); not inside a stream capture on first use
cudaError_t jit_kernel(JitKernel* k, int device, int smem_limit, const void** fn);
unsigned long long jit_ready_count(); // entries that have become ready so far (a captured graph is stale when this moves)
void jit_counters(unsigned long long* compiled, unsigned long long* disk_hits, unsigned long long* failed, double* compile_ms);
} // namespace dmb
