// Instantiates the solver kernels for the generated model "jerk_dense" (FP64 and FP32).
#include "bnmpc_kernels.cuh"
BNMPC_DEFINE_MODEL_OPS(bnmpc::Model_jerk_dense, bnmpc::KIND_JERK, ops_jerk_dense)
