// Evaluation metrics of the score consumer (reference core/utils/metric.py:132-157, Metric._get_correct_score):
// top-k predictions per clip, correctness flags against the target and the confusion-matrix update, in one launch with
// no host round trip (the reference walks the batch in a Python loop that indexes device tensors element by element).
#include "common.cuh"

namespace tbn {

constexpr int TOPK_MAX = 16;

// One warp per clip.  k passes of a warp arg-max over the scores not selected yet; ties go to the lowest class index
// (torch.topk is unspecified there).  correct[j][i] = (j-th best class of clip i == target[i]);
// conf[target[i]][best class of clip i] += 1.
__global__ void topk_confusion_kernel(const float* __restrict__ scores, const long long* __restrict__ target, int b, int c,
                                      int maxk, unsigned char* __restrict__ correct, long long* __restrict__ preds,
                                      int* __restrict__ rank, float* __restrict__ conf) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= b) return;
    const float* row = scores + (size_t)warp * c;
    const long long tgt = target[warp];
    int chosen[TOPK_MAX];
    int hit = maxk;  // position of the target among the maxk best classes (maxk = not among them)
    for (int j = 0; j < maxk; ++j) {
        float best = -INFINITY;
        int best_i = 0x7fffffff;
        for (int i = lane; i < c; i += 32) {
            bool taken = false;
            for (int t = 0; t < j; ++t) taken |= (chosen[t] == i);
            const float v = row[i];
            // NaN never wins; -inf scores are still selectable (best_i starts invalid)
            if (!taken && (v > best || (best_i == 0x7fffffff && v == best))) { best = v; best_i = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (oi != 0x7fffffff && (best_i == 0x7fffffff || ov > best || (ov == best && oi < best_i))) { best = ov; best_i = oi; }
        }
        chosen[j] = best_i;
        if (best_i == (int)tgt && hit == maxk) hit = j;
        if (lane == 0) {
            correct[(size_t)j * b + warp] = (unsigned char)(best_i == (int)tgt);
            if (preds) preds[(size_t)j * b + warp] = best_i;
            if (j == 0 && conf && tgt >= 0 && tgt < c && best_i < c) atomicAdd(conf + (size_t)tgt * c + best_i, 1.0f);
        }
    }
    if (lane == 0 && rank) rank[warp] = hit;
}

}  // namespace tbn

using namespace tbn;

extern "C" int tbn_topk_confusion(const float* scores, const long long* target, int b, int c, int maxk, unsigned char* correct,
                                  long long* preds, int* rank, float* conf_mat, void* stream) {
    TBN_CHECK_ARG(scores && target && correct, "topk_confusion: null pointer");
    TBN_CHECK_ARG(b > 0 && c > 0, "topk_confusion: bad shape (%d, %d)", b, c);
    TBN_CHECK_ARG(maxk >= 1 && maxk <= TOPK_MAX && maxk <= c, "topk_confusion: maxk must be in 1..min(%d, classes) (got %d)", TOPK_MAX,
                  maxk);
    const int warps_per_block = 8;
    topk_confusion_kernel<<<cdiv(b, warps_per_block), warps_per_block * 32, 0, as_stream(stream)>>>(scores, target, b, c, maxk, correct,
                                                                                                     preds, rank, conf_mat);
    TBN_LAUNCH_CHECK("topk_confusion launch");
    return 0;
}
