// c_api.cu — library identification, status strings and process-global tuning knobs.
#include <string.h>

#include "common.cuh"

namespace {
struct Knob { const char* key; int value; };
Knob g_knobs[] = {{"ctc_k", 0}, {"ctc_lin", 1}, {"ctc_ws", 1}, {"beam_fast", 1}, {"beam_two_phase", 1}, {"ctc_grad_warps", 0}, {"gemm_dbg", 0}, {"pdl", 1}, {"ctc_pf", 1}, {"beam_pf", 1}, {"lstm_tag", 1}, {"lstm_groups", 0}, {"ctc_stage", 1}, {"ctc_overlap", 1}, {"ctc_stamp", 0}, {"beam_fused", -1}, {"beam_fused_grid", 0}};
}  // namespace

int avctc_tuning_get(const char* key, int dflt) {
    for (auto& k : g_knobs)
        if (strcmp(k.key, key) == 0) return k.value;
    return dflt;
}

extern "C" int avctc_set_tuning(const char* key, int value) {
    if (!key) return AVCTC_ERR_BAD_ARG;
    for (auto& k : g_knobs)
        if (strcmp(k.key, key) == 0) { k.value = value; return AVCTC_OK; }
    return AVCTC_ERR_BAD_ARG;
}

extern "C" const char* avctc_version(void) { return "avctc_b200 0.1 sm_100a"; }

extern "C" const char* avctc_status_string(int status) {
    switch (status) {
        case AVCTC_OK: return "ok";
        case AVCTC_ERR_BAD_ARG: return "bad argument (null pointer, negative size or bad enum)";
        case AVCTC_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
        case AVCTC_ERR_WORKSPACE: return "workspace too small";
        case AVCTC_ERR_ALIGNMENT: return "pointer or stride alignment requirement violated";
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown avctc status";
}
