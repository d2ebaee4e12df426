// export.cu -- reference-layout view of the tree: KDNode[] records and particle offsets exactly as
// KDTree.serialize() returns them (pynbody/kdtree/__init__.py:17-30,136-163; kd.h:29-40).
//
// The device tree the kernels traverse uses fixed 32-particle buckets; the reference's tree has
// median-split ranges and a user-chosen leafsize (kd.cpp:98-111,176-199).  When a caller asks for the
// serialized form, a second, reference-shaped tree is built on the GPU from the resident particles
// (same build code, TreeShape::packed = 0), downloaded, and discarded.
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
}  // namespace

void export_reference_tree(Tree &t, void *kdnodes_out, int64_t n_nodes, int64_t *offsets_out)
{
    const TreeShape shape = reference_shape(t.n, t.user_leafsize);
    const int64_t nsplit = (int64_t)1 << shape.levels;
    if (kdnodes_out && n_nodes != 2 * nsplit) throw Error(PNB_ERR_VALUE, "kdnodes array has the wrong length");
    const size_t ts = tsize(t.dtype);
    // file-order coordinate columns from the resident tree-order particles
    DevBuf cols(ts * 3 * t.n);
    permute_to_file(t, t.x.p, cols.p, 1, (int)ts, t.stream);
    permute_to_file(t, t.y.p, (char *)cols.p + ts * t.n, 1, (int)ts, t.stream);
    permute_to_file(t, t.z.p, (char *)cols.p + 2 * ts * t.n, 1, (int)ts, t.stream);
    Tree ref;
    ref.dtype = t.dtype; ref.n = t.n; ref.user_leafsize = t.user_leafsize; ref.stream = t.stream;
    build_tree(ref, cols, nullptr, shape, /*want_soa=*/false);

    if (offsets_out) {
        std::vector<uint32_t> o((size_t)t.n);
        PNB_CUDA(cudaMemcpy(o.data(), ref.order.p, sizeof(uint32_t) * (size_t)t.n, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < t.n; ++i) offsets_out[i] = (int64_t)o[(size_t)i];
    }
    if (!kdnodes_out) return;
    const size_t nn = (size_t)2 * nsplit;
    std::vector<double> box(nn * 6), split(nn);
    std::vector<int8_t> dim(nn);
    if (t.dtype == PNB_F64) {
        PNB_CUDA(cudaMemcpy(box.data(), ref.box.p, sizeof(double) * nn * 6, cudaMemcpyDeviceToHost));
        PNB_CUDA(cudaMemcpy(split.data(), ref.split_val.p, sizeof(double) * nn, cudaMemcpyDeviceToHost));
    } else {
        std::vector<float> b(nn * 6), s(nn);
        PNB_CUDA(cudaMemcpy(b.data(), ref.box.p, sizeof(float) * nn * 6, cudaMemcpyDeviceToHost));
        PNB_CUDA(cudaMemcpy(s.data(), ref.split_val.p, sizeof(float) * nn, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < nn * 6; ++i) box[i] = b[i];
        for (size_t i = 0; i < nn; ++i) split[i] = s[i];
    }
    PNB_CUDA(cudaMemcpy(dim.data(), ref.split_dim.p, nn, cudaMemcpyDeviceToHost));
    RefKDNode *o = reinterpret_cast<RefKDNode *>(kdnodes_out);
    memset(o, 0, sizeof(RefKDNode) * nn);
    for (int level = 0; level <= shape.levels; ++level) {
        const int64_t cnt = (int64_t)1 << level;
        for (int64_t j = 0; j < cnt; ++j) {
            const int64_t node = cnt + j;
            RefKDNode &r = o[node];
            shape_seg_of_node(shape, level, j, r.pLower, r.pUpper);
            for (int d = 0; d < 3; ++d) { r.fMin[d] = box[node * 6 + d]; r.fMax[d] = box[node * 6 + 3 + d]; }
            if (level == shape.levels) { r.iDim = -1; r.fSplit = 0; }         // kd.cpp:158-166: bucket
            else { r.iDim = dim[node]; r.fSplit = split[node]; }
        }
    }
}

}  // namespace pnb
