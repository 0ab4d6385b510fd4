// Probe: contract-M (PARAM-sourced chain, mix only) through process_host vs process_device variants.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <chrono>
#include "mlb200.h"
static double now(){ return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
  const int V = 65536, T = 64;
  mlb_init(0);
  const mlb_node nodes[5] = {{MLB_OP_PARAM, {-1, -1, -1}, 0}, {MLB_OP_SINE, {0, -1, -1}, 0}, {MLB_OP_LOPASS, {1, -1, -1}, 0},
                             {MLB_OP_PARAM, {-1, -1, -1}, 0}, {MLB_OP_MULTIPLY, {2, 3, -1}, 0}};
  const int32_t outs[1] = {4};
  mlb_graph* g = nullptr;
  if (mlb_graph_create(nodes, 5, outs, 1, V, MLB_GRAPH_EXACT, &g) != MLB_OK) { printf("create: %s\n", mlb_last_error()); return 1; }
  mlb_layout L; mlb_graph_layout_of(g, &L);
  printf("kernel %s n_state %d n_coef %d\n", mlb_graph_kernel_name(g), L.n_state_words, L.n_coef_words);
  std::vector<float> coef((size_t)L.n_coef_words * V);
  float c[3]; mlb_coeffs_lopass(0.1f, 0.5f, c);
  for (int v = 0; v < V; ++v) { coef[v] = 0.001f + 1e-7f * v; coef[V + v] = c[0]; coef[2 * V + v] = c[1]; coef[3 * V + v] = c[2]; coef[4 * V + v] = 0.5f; }
  mlb_graph_set_coefs(g, coef.data());
  float* h_mix; cudaMallocHost(&h_mix, T * 64 * 4);
  float* d_mix; cudaMalloc(&d_mix, T * 64 * 4);
  float ms;
  for (int rep = 0; rep < 2; ++rep)
  {
    for (int i = 0; i < 3; ++i) mlb_graph_process_host(g, nullptr, nullptr, h_mix, T);
    double t0 = now();
    for (int i = 0; i < 20; ++i) mlb_graph_process_host(g, nullptr, nullptr, h_mix, T);
    mlb_graph_last_kernel_ms(g, &ms);
    printf("process_host: %.3f ms/call kernel %.3f\n", (now() - t0) / 20 * 1e3, ms);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int i = 0; i < 3; ++i) { mlb_graph_process_device(g, nullptr, nullptr, d_mix, T, s); cudaStreamSynchronize(s); }
    t0 = now();
    for (int i = 0; i < 20; ++i) { mlb_graph_process_device(g, nullptr, nullptr, d_mix, T, s); cudaStreamSynchronize(s); }
    mlb_graph_last_kernel_ms(g, &ms);
    printf("process_device own nonblocking stream + stream sync: %.3f ms/call kernel %.3f\n", (now() - t0) / 20 * 1e3, ms);
    t0 = now();
    for (int i = 0; i < 20; ++i) { mlb_graph_process_device(g, nullptr, nullptr, d_mix, T, s); cudaMemcpyAsync(h_mix, d_mix, T * 64 * 4, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); }
    mlb_graph_last_kernel_ms(g, &ms);
    printf("  ... + D2H mix: %.3f ms/call kernel %.3f\n", (now() - t0) / 20 * 1e3, ms);
    t0 = now();
    for (int i = 0; i < 20; ++i) { mlb_graph_process_device(g, nullptr, nullptr, d_mix, T, nullptr); cudaDeviceSynchronize(); }
    mlb_graph_last_kernel_ms(g, &ms);
    printf("process_device legacy stream: %.3f ms/call kernel %.3f\n", (now() - t0) / 20 * 1e3, ms);
  }
  mlb_graph_destroy(g);
  return 0;
}
