// export.cu -- reference-layout view of the device tree: KDNode[] records as KDTree.serialize()
// returns them (pynbody/kdtree/__init__.py:17-30,136-163; kd.h:29-40).
#include "common.cuh"

namespace pnb {

namespace {
#pragma pack(push, 1)
struct RefKDNode {            // kd.h:29-40, 80 bytes
    double fSplit;
    double fMin[3], fMax[3];
    int32_t iDim;
    int32_t pad;
    int64_t pLower, pUpper;
};
#pragma pack(pop)
static_assert(sizeof(RefKDNode) == 80, "KDNode layout");

inline void seg_of_node_host(int64_t n, int level, int64_t j, int64_t &lo, int64_t &hi)
{
    lo = 0; hi = n - 1;
    for (int b = level - 1; b >= 0; --b) {
        int64_t m = (lo + hi) >> 1;
        if ((j >> b) & 1) lo = m + 1; else hi = m;
    }
}
}  // namespace

void export_kdnodes(Tree &t, void *out, int64_t n_nodes)
{
    int64_t n = t.n, nsplit_user = 1;
    int levels_user = 0;
    while (n > t.user_leafsize) { n >>= 1; nsplit_user <<= 1; ++levels_user; }     // kd.cpp:98-111
    if (n_nodes != 2 * nsplit_user) throw Error(PNB_ERR_VALUE, "kdnodes array has the wrong length");
    if (levels_user > t.levels)
        throw Error(PNB_ERR_VALUE, "reference-layout export needs leafsize >= 32 for this particle count "
                                   "(the device tree uses 32-particle buckets)");
    const size_t nn = (size_t)2 * nsplit_user;
    std::vector<double> box(nn * 6), split(nn);
    std::vector<int8_t> dim(nn);
    if (t.dtype == PNB_F64) {
        PNB_CUDA(cudaMemcpy(box.data(), t.box.p, sizeof(double) * nn * 6, cudaMemcpyDeviceToHost));
        PNB_CUDA(cudaMemcpy(split.data(), t.split_val.p, sizeof(double) * nn, cudaMemcpyDeviceToHost));
    } else {
        std::vector<float> b(nn * 6), s(nn);
        PNB_CUDA(cudaMemcpy(b.data(), t.box.p, sizeof(float) * nn * 6, cudaMemcpyDeviceToHost));
        PNB_CUDA(cudaMemcpy(s.data(), t.split_val.p, sizeof(float) * nn, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < nn * 6; ++i) box[i] = b[i];
        for (size_t i = 0; i < nn; ++i) split[i] = s[i];
    }
    PNB_CUDA(cudaMemcpy(dim.data(), t.split_dim.p, nn, cudaMemcpyDeviceToHost));
    RefKDNode *o = reinterpret_cast<RefKDNode *>(out);
    memset(o, 0, sizeof(RefKDNode) * nn);
    for (int level = 0; level <= levels_user; ++level) {
        const int64_t cnt = (int64_t)1 << level;
        for (int64_t j = 0; j < cnt; ++j) {
            const int64_t node = cnt + j;
            RefKDNode &r = o[node];
            seg_of_node_host(t.n, level, j, r.pLower, r.pUpper);
            for (int d = 0; d < 3; ++d) { r.fMin[d] = box[node * 6 + d]; r.fMax[d] = box[node * 6 + 3 + d]; }
            if (level == levels_user) { r.iDim = -1; r.fSplit = 0; }
            else { r.iDim = dim[node]; r.fSplit = split[node]; }
        }
    }
}

}  // namespace pnb
