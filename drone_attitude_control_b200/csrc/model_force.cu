// Instantiates the solver kernels for the generated model "force" (FP64 and FP32).
#include "bnmpc_kernels.cuh"
BNMPC_DEFINE_MODEL_OPS(bnmpc::Model_force, bnmpc::KIND_FORCE, ops_force, true)
