// Instantiates the solver kernels for the generated model "plant" used as controller model (BNMPC_MODEL_THRUST): the
// general path - one block, sensitivities recomputed at every stage and SQP iteration.
#include "bnmpc_kernels.cuh"
BNMPC_DEFINE_MODEL_OPS(bnmpc::Model_plant, bnmpc::KIND_THRUST, ops_thrust, false)
