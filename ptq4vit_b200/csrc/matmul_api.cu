// C-ABI for the head-wise MatMul scale-factor search (placeholder until the planner lands).
#include "../../include/ptq4vit_b200.h"
#include "prep.cuh"

extern "C" int p4v_matmul_workspace_bytes(const p4v_matmul_desc*, size_t*) { p4v_set_error("matmul: not built yet"); return 1; }
extern "C" int p4v_matmul_score_log_floats(const p4v_matmul_desc*, size_t*) { p4v_set_error("matmul: not built yet"); return 1; }
extern "C" int p4v_matmul_calibrate(const p4v_matmul_desc*, const float*, const float*, const float*, const float*, void*, size_t,
                                    float*, float*, float*, float*, void*) { p4v_set_error("matmul: not built yet"); return 1; }
extern "C" int p4v_matmul_quant_forward_workspace_bytes(const p4v_matmul_desc*, size_t*) { p4v_set_error("matmul: not built yet"); return 1; }
extern "C" int p4v_matmul_quant_forward(const p4v_matmul_desc*, const float*, const float*, const float*, const float*, const float*,
                                        void*, size_t, float*, void*) { p4v_set_error("matmul: not built yet"); return 1; }
