// The whole C-ABI library compiled for the host SIMT interpreter -- TEST INFRASTRUCTURE.
//   g++ -std=c++17 -O1 -fPIC -shared -ffp-contract=off -I tests/simt/stub -I torchmd_b200/csrc -o tests/simt/libtmd_simt.so tests/simt/simt_lib.cpp
// tests/simt/stub/cuda_runtime.h shadows the CUDA header: kernels become plain functions, launches
// run in simt.h's interpreter, "device" memory is host memory.  tests/test_simt_kernels.py drives
// the C ABI with numpy arrays to check the kernels' logic on the CPU.  tmd_version() returns -100
// here, and torchmd_b200/_lib.py refuses to load a library that says so: this is not a CPU path of
// the product.
#define TMD_SIMT_HOST 1
#include "../../torchmd_b200/csrc/tmd_b200.cu"

// what the record-and-replay graph model did (tests/test_simt_kernels.py): replays, IF bodies run / skipped
extern "C" void simt_graph_counters(long long* out) {
  out[0] = simt_stub::state().graph_launches;
  out[1] = simt_stub::state().bodies_run;
  out[2] = simt_stub::state().bodies_skipped;
}

// stream capture driven from Python (the tests' stand-in for torch.cuda.graph around library calls)
extern "C" int simt_capture_begin(void* stream) { return (int)cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeGlobal); }
extern "C" void* simt_capture_end(void* stream) {
  cudaGraph_t g = nullptr;
  return cudaStreamEndCapture((cudaStream_t)stream, &g) == cudaSuccess ? (void*)g : nullptr;
}
extern "C" int simt_graph_replay(void* graph) { return graph ? (int)cudaGraphLaunch((cudaGraphExec_t)graph, nullptr) : -1; }
