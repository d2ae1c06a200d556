// Tensor-core (tcgen05) matcher -- placeholder until the TMEM kernel lands; reports unsupported so
// callers asking for algo=1 fail loudly instead of silently taking another path.
#include "kb_common.cuh"

size_t kb_match_tc_workspace_bytes(int B, int n_max, int m_max, int D) {
    (void)B; (void)n_max; (void)m_max; (void)D;
    return 256;
}

int kb_match_tc_run(const float* d0, const float* d1, const int* n0, const int* n1, int B, int n_max, int m_max,
                    int D, double max_distance, int cross_check, int* pairs, double* dist, int* count, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
    (void)d0; (void)d1; (void)n0; (void)n1; (void)B; (void)n_max; (void)m_max; (void)D; (void)max_distance;
    (void)cross_check; (void)pairs; (void)dist; (void)count; (void)ws; (void)ws_bytes; (void)st;
    return KB_ERR_UNSUPPORTED;
}
