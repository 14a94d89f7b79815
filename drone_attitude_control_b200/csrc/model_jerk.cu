// Instantiates the solver kernels for the generated model "jerk" (FP64 and FP32).
#include "bnmpc_kernels.cuh"
BNMPC_DEFINE_MODEL_OPS(bnmpc::Model_jerk, bnmpc::KIND_JERK, ops_jerk, true)
