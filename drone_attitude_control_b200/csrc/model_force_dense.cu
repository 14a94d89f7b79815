// Instantiates the solver kernels for the generated model "force_dense" (FP64 and FP32).
#include "bnmpc_kernels.cuh"
BNMPC_DEFINE_MODEL_OPS(bnmpc::Model_force_dense, bnmpc::KIND_FORCE, ops_force_dense, false)
