#include "common.cuh"

unsigned long long g_cp_launches = 0;

extern "C" int cp_version(void) { return 110; }   // 0.1.1: + l2, confusion matrix, preprocessing, CP_ENGINE_TC_FP16

// number of kernels this library has launched in this process (host-side counter, not thread-safe
// across concurrent callers; used by bench.py's gpu_launches)
extern "C" unsigned long long cp_launch_count(void) { return g_cp_launches; }

extern "C" const char* cp_status_string(int status) {
    switch (status) {
        case CP_OK: return "ok";
        case CP_ERR_ARG: return "invalid argument (null pointer, bad size or misaligned workspace)";
        case CP_ERR_WORKSPACE: return "workspace too small";
        case CP_ERR_UNSUPPORTED: return "unsupported mode for this entry point";
        case CP_ERR_COLLECTIVE: return "the caller's all-reduce callback (SyncBN) reported a failure";
        default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown libcpros status";
    }
}
