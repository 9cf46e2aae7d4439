// trainer_stress.cpp -- AddressSanitizer + UBSan run of the host layer's graph executor set-up and snapshot code
// (tests/test_trainer_sim.py builds and runs it).  TEST INFRASTRUCTURE ONLY.  For every net prototxt given on the command line:
// build the TrainNet through the C handle API (layers, parameter arena, fusion pass, bucket plan), write every parameter and its
// momentum history, Snapshot, build a second trainer, Restore, compare; then CopyTrainedLayersFrom.  Nets named *_step.prototxt also
// run forward / backward and three Solver::Step iterations on the host stand-ins for the kernels.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

extern "C" {
const char* b2h_last_error();
void* b2h_trainer_create(const char* net_src, int net_is_text, const char* solver_src, int solver_is_text, int batch, int num_classes,
                         unsigned long long seed, int math, int def_channels, int def_size);
void b2h_trainer_destroy(void*);
int b2h_trainer_num_params(void*);
int b2h_trainer_num_learnable(void*);
long long b2h_trainer_param_count(void*, int);
int b2h_trainer_param(void*, int i, int what, int set, float* buf);
int b2h_trainer_snapshot(void*, const char* prefix, char* state_path, int len);
int b2h_trainer_restore(void*, const char* state_path);
int b2h_trainer_copy_from(void*, const char* model_path, int* copied);
long long b2h_trainer_arena_floats(void*);
int b2h_trainer_step(void*, int nsteps, int copy_input);
int b2h_trainer_forward_backward(void*, float* loss);
int b2h_trainer_clear_param_diffs(void*);
int b2h_trainer_loss(void*, float* loss);
}
#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "trainer_stress: %s:%d: %s  [%s]\n", __FILE__, __LINE__, #c, b2h_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: trainer_stress <scratch dir> <channels> <size> net.prototxt...\n"); return 2; }
  setenv("B2C_NCCL_ARENA", "0", 1);
  const std::string dir = argv[1];
  const int ch = atoi(argv[2]), sz = atoi(argv[3]);
  const char* solver = "base_lr: 0.05 lr_policy: \"step\" stepsize: 10 gamma: 0.5 momentum: 0.9 weight_decay: 0.0005 max_iter: 100";
  for (int a = 4; a < argc; ++a) {
    void* t = b2h_trainer_create(argv[a], 0, solver, 1, 2, 1000, 1701, 0, ch, sz);
    REQUIRE(t != nullptr);
    const int n = b2h_trainer_num_params(t);
    REQUIRE(n > 0 && b2h_trainer_num_learnable(t) >= n && b2h_trainer_arena_floats(t) > 0);
    std::vector<std::vector<float>> vals;
    for (int i = 0; i < n; ++i)
      for (int what : {0, 2}) {
        std::vector<float> v((size_t)b2h_trainer_param_count(t, i));
        for (size_t k = 0; k < v.size(); ++k) v[k] = (float)((i * 131 + what * 17 + k) % 1009) * 0.001f - 0.5f;
        REQUIRE(b2h_trainer_param(t, i, what, 1, v.data()) == 0);
        vals.push_back(v);
      }
    if (std::string(argv[a]).find("_step.prototxt") != std::string::npos) {
      // a small net: whole iterations on the host stand-ins for the kernels (fake_kernels*.cpp), so that the executor's own
      // allocations -- shadow diffs, fused-layer scratch, the update's segment tables -- are exercised under the sanitizers
      float loss = 0.f;
      REQUIRE(b2h_trainer_forward_backward(t, &loss) == 0 && loss == loss);
      REQUIRE(b2h_trainer_clear_param_diffs(t) == 0);
      REQUIRE(b2h_trainer_step(t, 3, 1) == 0);
      REQUIRE(b2h_trainer_loss(t, &loss) == 0 && loss == loss);
      size_t k = 0;                                        // the snapshot below must hold the TRAINED state
      for (int i = 0; i < n; ++i)
        for (int what : {0, 2}) { REQUIRE(b2h_trainer_param(t, i, what, 0, vals[k].data()) == 0); ++k; }
    }
    char state[1024];
    const std::string prefix = dir + "/net" + std::to_string(a);
    REQUIRE(b2h_trainer_snapshot(t, prefix.c_str(), state, sizeof(state)) == 0);
    void* u = b2h_trainer_create(argv[a], 0, solver, 1, 2, 1000, 7, 0, ch, sz);
    REQUIRE(u != nullptr);
    REQUIRE(b2h_trainer_restore(u, state) == 0);
    size_t k = 0;
    for (int i = 0; i < n; ++i)
      for (int what : {0, 2}) {
        std::vector<float> v(vals[k].size());
        REQUIRE(b2h_trainer_param(u, i, what, 0, v.data()) == 0);
        REQUIRE(v == vals[k]);
        ++k;
      }
    int copied = 0;
    void* w = b2h_trainer_create(argv[a], 0, solver, 1, 2, 1000, 9, 0, ch, sz);
    std::string model = state;                              // <prefix>_iter_<N>.solverstate -> .caffemodel
    model.replace(model.rfind(".solverstate"), std::string::npos, ".caffemodel");
    REQUIRE(w != nullptr && b2h_trainer_copy_from(w, model.c_str(), &copied) == 0 && copied > 0);
    b2h_trainer_destroy(w);
    b2h_trainer_destroy(u);
    b2h_trainer_destroy(t);
    printf("trainer_stress: %s ok (%d trainable blobs)\n", argv[a], n);
  }
  printf("trainer_stress ok\n");
  return 0;
}
