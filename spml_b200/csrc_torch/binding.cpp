// Host-side binding of the stage-group C ABI (include/spml_b200.h) for PyTorch callers.
//
// libspml_b200.so knows nothing about torch: it takes device pointers, sizes and a stream.
// What a caller has to do around each call is allocate the outputs, fill the argument struct
// and wrap the results as tensors.  Done in Python (spml_b200/ops.py over ctypes) that
// bookkeeping costs more host time than the GPU needs for the work (profiles/
// r2_step_profile_b1.txt: ~0.5 ms of a 1.2 ms step), and the step has one host
// synchronisation, so the host IS the critical path.  This file is the same bookkeeping in
// C++ (ATen allocations, no arithmetic); spml_b200/ops.py uses it when it is built and the
// ctypes path otherwise.  Both end in the same library calls.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <optional>
#include <vector>

#include "../../include/spml_b200.h"

namespace {

using at::Tensor;
using OptTensor = std::optional<Tensor>;

void check(int rc, const char* what) {
  TORCH_CHECK(rc == SPML_OK, what, " failed (", rc, "): ", spml_last_error());
}

void* stream_of(const Tensor& t) {
  return at::cuda::getCurrentCUDAStream(t.get_device()).stream();
}

Tensor f32c(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, ": spml_b200 needs CUDA tensors (no CPU path)");
  TORCH_CHECK(t.scalar_type() == at::kFloat, name, ": expected float32");
  return t.is_contiguous() ? t : t.contiguous();
}

Tensor i64c(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, ": spml_b200 needs CUDA tensors (no CPU path)");
  Tensor r = t.scalar_type() == at::kLong ? t : t.to(at::kLong);
  return r.is_contiguous() ? r : r.contiguous();
}

template <typename T>
T* ptr_or_null(const OptTensor& t) {
  return t.has_value() && t->defined() ? t->data_ptr<T>() : nullptr;
}

Tensor workspace(size_t bytes, const Tensor& like) {
  return at::empty({(int64_t)std::max<size_t>(bytes, 16)}, like.options().dtype(at::kByte));
}

Tensor rows_view(const Tensor& buf, int64_t rows, int64_t cols, int64_t offset) {
  if (cols == 0) return at::as_strided(buf, {rows}, {1}, offset);
  return at::as_strided(buf, {rows, cols}, {cols, 1}, offset);
}

// ------------------------------------------------------------------------------------ A8

// Returns {e, el, labels, cluster_index, batch_index, [sem, inst,] fbuf, ibuf} and the sizes.
std::tuple<std::vector<Tensor>, int64_t, int64_t, int64_t> segment_fwd(
    const Tensor& emb_in, const OptTensor& loc_in, const OptTensor& labels_in,
    const OptTensor& sem_in, const OptTensor& inst_in, int64_t divisor, int64_t semantic_ignore,
    bool has_ignore, int64_t ignore_index, const OptTensor& ignore_dev_in, const Tensor& seeds_in,
    const OptTensor& k_per_image, int64_t num_k, int64_t iterations, int64_t batch_index_offset,
    const Tensor& status) {
  const Tensor emb = f32c(emb_in, "segment_by_kmeans(embeddings)");
  TORCH_CHECK(emb.dim() == 4, "embeddings must be [batch, channels, height, width]");
  c10::cuda::CUDAGuard guard(emb.device());
  const int64_t B = emb.size(0), D = emb.size(1), H = emb.size(2), W = emb.size(3);
  const int64_t n = H * W, cap = B * n;
  spml_cluster_args a{};
  Tensor loc;
  int64_t loc_ch = 0;
  if (loc_in.has_value() && loc_in->defined()) {
    const Tensor& l = *loc_in;
    TORCH_CHECK(l.dim() == 4 && l.size(1) == H && l.size(2) == W,
                "local_features must be [batch, H, W, C]");
    loc_ch = l.size(3);
    if (l.stride(0) == 0 || l.size(0) == 1) {
      loc = f32c(l.select(0, 0), "local_features");
    } else {
      loc = f32c(l, "local_features");
      a.loc_batch_stride = n * loc_ch;
    }
    a.loc = loc.data_ptr<float>();
  }
  const int64_t DL = D + loc_ch;
  const Tensor seeds = i64c(seeds_in, "cluster_indices");
  Tensor labels, sem, inst, ignore_dev;
  if (labels_in.has_value() && labels_in->defined()) {
    labels = i64c(*labels_in, "segment_by_kmeans(labels)");
    a.labels = labels.data_ptr<int64_t>();
    a.has_ignore = has_ignore ? 1 : 0;
    a.ignore_index = ignore_index;
    if (ignore_dev_in.has_value() && ignore_dev_in->defined()) {
      ignore_dev = i64c(ignore_dev_in->reshape({1}), "ignore_index");
      a.ignore_index_dev = ignore_dev.data_ptr<int64_t>();
    }
  } else {
    TORCH_CHECK(sem_in.has_value() && inst_in.has_value(), "labels or semantic + instance labels");
    sem = i64c(*sem_in, "semantic_labels");
    inst = i64c(*inst_in, "instance_labels");
    a.sem = sem.data_ptr<int64_t>();
    a.inst = inst.data_ptr<int64_t>();
    a.semantic_ignore = semantic_ignore;
  }
  a.label_divisor = divisor;
  a.emb = emb.data_ptr<float>();
  a.seeds = seeds.data_ptr<int64_t>();
  a.seed_batch_stride = seeds.dim() == 2 ? 0 : n;
  a.k_per_image = ptr_or_null<int32_t>(k_per_image);
  a.batch_index_offset = batch_index_offset;
  a.batch = (int32_t)B, a.dim = (int32_t)D, a.n = (int32_t)n, a.loc_ch = (int32_t)loc_ch;
  a.num_clusters = (int32_t)num_k, a.iterations = (int32_t)iterations;
  a.eps = 1e-12f;

  const auto fopt = emb.options();
  Tensor fbuf = at::empty({cap * (D + DL + 2)}, fopt);
  Tensor lbuf = at::empty({cap * 5}, fopt.dtype(at::kLong));
  const int64_t head = (B + 2 + 4 + 3) / 4 * 4;
  Tensor ibuf = at::empty({head + cap * 3}, fopt.dtype(at::kInt));
  const size_t ws_bytes =
      spml_segment_by_kmeans_workspace_bytes((int)B, (int)n, (int)DL, (int)num_k, (int)iterations);
  Tensor ws = workspace(ws_bytes, emb);
  float* fp = fbuf.data_ptr<float>();
  int64_t* lp = lbuf.data_ptr<int64_t>();
  int32_t* ip = ibuf.data_ptr<int32_t>();
  a.e = fp, a.el = fp + cap * D;
  a.nx = fp + cap * (D + DL), a.nc = fp + cap * (D + DL + 1);
  a.labels_out = lp, a.batch_out = lp + cap, a.segment_ids = lp + 2 * cap;
  if (divisor > 0) a.sem_out = lp + 3 * cap, a.inst_out = lp + 4 * cap;
  a.img_off = ip, a.num_segments = ip + (B + 1);
  a.counts_dev = ip + (B + 2);
  a.dst = ip + head, a.kmeans_labels = ip + head + cap, a.seed_out = ip + head + 2 * cap;
  int32_t counts[4] = {0, 0, 0, 0};
  a.counts_host = counts;
  a.status = status.data_ptr<int32_t>();
  {
    pybind11::gil_scoped_release release;     // the call blocks on the step's one host sync
    check(spml_segment_by_kmeans(&a, ws.data_ptr(), ws_bytes, stream_of(emb)),
          "spml_segment_by_kmeans");
  }
  const int64_t rows = counts[0], segments = counts[1];
  std::vector<Tensor> out;
  out.reserve(9);
  out.push_back(rows_view(fbuf, rows, D, 0));
  out.push_back(rows_view(fbuf, rows, DL, cap * D));
  out.push_back(rows_view(lbuf, rows, 0, 0));           // labels
  out.push_back(rows_view(lbuf, rows, 0, 2 * cap));     // cluster_index
  out.push_back(rows_view(lbuf, rows, 0, cap));         // batch_index
  if (divisor > 0) {
    out.push_back(rows_view(lbuf, rows, 0, 3 * cap));
    out.push_back(rows_view(lbuf, rows, 0, 4 * cap));
  }
  out.push_back(fbuf);
  out.push_back(ibuf);
  return {out, rows, segments, (int64_t)counts[2]};
}

Tensor segment_bwd(const Tensor& fbuf, const Tensor& ibuf, const OptTensor& de_in,
                   const OptTensor& del_in, int64_t B, int64_t D, int64_t loc_ch, int64_t H,
                   int64_t W, int64_t head) {
  c10::cuda::CUDAGuard guard(fbuf.device());
  const int64_t cap = B * H * W, DL = D + loc_ch;
  Tensor de, del;
  if (de_in.has_value() && de_in->defined()) de = f32c(*de_in, "d(cluster_embedding)");
  if (del_in.has_value() && del_in->defined()) del = f32c(*del_in, "d(cluster_embedding_with_loc)");
  Tensor demb = at::empty({B, D, H, W}, fbuf.options());
  const float* fp = fbuf.data_ptr<float>();
  check(spml_normalize_pack_bwd(de.defined() ? de.data_ptr<float>() : nullptr,
                                del.defined() ? del.data_ptr<float>() : nullptr, fp, fp + cap * D,
                                fp + cap * (D + DL), fp + cap * (D + DL + 1),
                                ibuf.data_ptr<int32_t>() + head, (int)B, (int)D, (int)loc_ch,
                                (int)(H * W), 1e-12f, demb.data_ptr<float>(), stream_of(fbuf)),
        "spml_normalize_pack_bwd");
  return demb;
}

// ------------------------------------------------------------------------------------ B1

std::vector<Tensor> gather_fwd(const Tensor& e_in, const Tensor& el_in, const Tensor& cid_in,
                               const Tensor& bid_in, const Tensor& sem_in, const Tensor& inst_in,
                               int64_t m, const Tensor& status) {
  const Tensor e = f32c(e_in, "gather(embeddings)"), el = f32c(el_in, "gather(embeddings_with_loc)");
  c10::cuda::CUDAGuard guard(e.device());
  const Tensor cid = i64c(cid_in, "cluster_indices"), bid = i64c(bid_in, "batch_indices");
  const Tensor sem = i64c(sem_in, "semantic_labels"), inst = i64c(inst_in, "instance_labels");
  const int64_t rows = e.size(0), D = e.size(1), DL = el.size(1);
  Tensor fbuf = at::empty({m * (D + DL + 2)}, e.options());
  Tensor plab = at::empty({3 * m}, e.options().dtype(at::kLong));
  const size_t ws_bytes = spml_gather_prototypes_workspace_bytes(m, (int)D, (int)DL);
  Tensor ws = workspace(ws_bytes, e);
  float* fp = fbuf.data_ptr<float>();
  int64_t* lp = plab.data_ptr<int64_t>();
  check(spml_gather_prototypes_fwd(
            e.data_ptr<float>(), el.data_ptr<float>(), rows, (int)D, (int)DL,
            cid.data_ptr<int64_t>(), bid.data_ptr<int64_t>(), sem.data_ptr<int64_t>(),
            inst.data_ptr<int64_t>(), m, 1e-12f, fp, fp + m * D, fp + m * (D + DL),
            fp + m * (D + DL + 1), lp, lp + m, lp + 2 * m, status.data_ptr<int32_t>(),
            ws.data_ptr(), ws_bytes, stream_of(e)),
        "spml_gather_prototypes_fwd");
  return {rows_view(fbuf, m, D, 0), rows_view(fbuf, m, DL, m * D), rows_view(plab, m, 0, 0),
          rows_view(plab, m, 0, m), rows_view(plab, m, 0, 2 * m), fbuf, cid};
}

std::vector<Tensor> gather_bwd(const Tensor& fbuf, const Tensor& cid, const OptTensor& dp_in,
                               const OptTensor& dpl_in, int64_t m, int64_t D, int64_t DL) {
  c10::cuda::CUDAGuard guard(fbuf.device());
  const int64_t rows = cid.size(0);
  Tensor dp, dpl, de, del;
  if (dp_in.has_value() && dp_in->defined()) {
    dp = f32c(*dp_in, "d(prototypes)");
    de = at::empty({rows, D}, fbuf.options());
  }
  if (dpl_in.has_value() && dpl_in->defined()) {
    dpl = f32c(*dpl_in, "d(prototypes_with_loc)");
    del = at::empty({rows, DL}, fbuf.options());
  }
  const float* fp = fbuf.data_ptr<float>();
  check(spml_gather_prototypes_bwd(
            dp.defined() ? dp.data_ptr<float>() : nullptr,
            dpl.defined() ? dpl.data_ptr<float>() : nullptr, fp, fp + m * D, fp + m * (D + DL),
            fp + m * (D + DL + 1), cid.data_ptr<int64_t>(), rows, (int)D, (int)DL, m, 1e-12f,
            de.defined() ? de.data_ptr<float>() : nullptr,
            del.defined() ? del.data_ptr<float>() : nullptr, stream_of(fbuf)),
        "spml_gather_prototypes_bwd");
  return {de, del};
}

// ------------------------------------------------------------------------------------ C4

// One forward / backward pair of Segsort*.losses(): owns the argument struct and keeps every
// tensor whose address is in it alive until the backward has been enqueued.
struct HeadCall : torch::CustomClassHolder {
  spml_head_args a{};
  std::vector<Tensor> keep;
  Tensor state, e, el, protos;

  const int64_t* i64p(const OptTensor& t, const char* name) {
    if (!t.has_value() || !t->defined()) return nullptr;
    keep.push_back(i64c(*t, name));
    return keep.back().data_ptr<int64_t>();
  }
  const float* f32p(const OptTensor& t, const char* name) {
    if (!t.has_value() || !t->defined()) return nullptr;
    keep.push_back(f32c(t->detach(), name));
    return keep.back().data_ptr<float>();
  }
  const int64_t* tags2d(const Tensor& t, const char* name, int64_t* ld, int64_t* rows) {
    Tensor r = i64c(t, name);
    if (r.dim() != 2) r = r.reshape({-1, r.size(-1)}).contiguous();
    keep.push_back(r);
    *ld = r.stride(0);
    if (rows) *rows = r.size(0);
    return r.data_ptr<int64_t>();
  }

  HeadCall(const Tensor& cid, const OptTensor& bid, const OptTensor& sem, const OptTensor& inst,
           const Tensor& psem, const OptTensor& pinst, const OptTensor& pbid, int64_t num_classes,
           int64_t enable, std::vector<double> kappas, std::vector<double> weights,
           int64_t max_groups, int64_t max_rows_per_group, const OptTensor& img_tags,
           const OptTensor& ptags, int64_t tag_col0, int64_t tag_col1,
           const std::vector<Tensor>& bank_protos, const std::vector<Tensor>& bank_sems,
           const std::vector<Tensor>& bank_bids, const std::vector<Tensor>& bank_tags,
           const std::vector<Tensor>& bank_locs, bool nn_tags, bool img_sim_on_plain,
           const OptTensor& protos_loc, double nn_threshold) {
    keep.reserve(16 + 4 * bank_protos.size());
    keep.push_back(i64c(cid.reshape({-1}), "cluster_index"));
    a.seg = keep.back().data_ptr<int64_t>();
    a.n = keep.back().size(0);
    a.bid = i64p(bid, "cluster_batch_index");
    a.sem = i64p(sem, "cluster_semantic_label");
    a.inst = i64p(inst, "cluster_instance_label");
    a.psem = i64p(psem, "prototype_semantic_label");
    a.pinst = i64p(pinst, "prototype_instance_label");
    a.pbid = i64p(pbid, "prototype_batch_index");
    a.m = psem.size(0);
    a.num_classes = num_classes;
    a.enable = (uint32_t)enable;
    a.kappa_ann = (float)kappas[0], a.kappa_occ = (float)kappas[1], a.kappa_sim = (float)kappas[2];
    a.weight_ann = (float)weights[0], a.weight_occ = (float)weights[1];
    a.weight_sim = (float)weights[2];
    a.max_groups = (int32_t)std::max<int64_t>(1, max_groups);
    a.max_rows_per_group = max_rows_per_group;
    a.nn_tags = nn_tags, a.img_sim_on_plain = img_sim_on_plain;
    a.nn_threshold = (float)nn_threshold;
    a.eps = 1e-12f;
    const bool occ = (enable & 2) != 0;
    if (occ && nn_tags) {
      a.protos_loc = f32p(protos_loc, "prototype_with_loc");
      a.dim_loc = (int32_t)keep.back().size(1);
      a.wide_tags = num_classes > 32;
    } else if (occ) {
      TORCH_CHECK(img_tags.has_value() && ptags.has_value(), "sem_occ needs the image tags");
      a.img_tags = tags2d(*img_tags, "semantic_tag", &a.img_tags_ld, &a.tag_rows);
      TORCH_CHECK(tag_col1 - tag_col0 <= 64 && tag_col1 <= keep.back().size(1),
                  "image tags: at most 64 tag columns inside the tag matrix");
      a.ptags = tags2d(*ptags, "prototype_semantic_tag", &a.ptags_ld, nullptr);
      a.tag_col0 = (int32_t)tag_col0, a.tag_col1 = (int32_t)tag_col1;
      a.wide_tags = tag_col1 - tag_col0 > 32;
    }
    TORCH_CHECK(bank_protos.size() <= SPML_MAX_BANK, "memory bank: at most ", SPML_MAX_BANK,
                " entries per call");
    a.num_bank = (int32_t)bank_protos.size();
    for (size_t i = 0; i < bank_protos.size(); ++i) {
      a.bank_protos[i] = f32p(bank_protos[i], "memory_prototype");
      a.bank_m[i] = bank_protos[i].size(0);
      a.bank_psem[i] = i64p(bank_sems[i], "memory_prototype_semantic_label");
      if (occ && nn_tags) {
        a.bank_protos_loc[i] = f32p(bank_locs[i], "memory_prototype_with_loc");
        a.bank_pbid[i] = i64p(bank_bids[i], "memory_prototype_batch_index");
      } else if (occ) {
        a.bank_tags[i] = tags2d(bank_tags[i], "memory_prototype_semantic_tag",
                                &a.bank_tags_ld[i], nullptr);
      }
    }
  }

  Tensor forward(const Tensor& e_in, const OptTensor& el_in, const Tensor& protos_in,
                 const Tensor& status) {
    e = f32c(e_in, "cluster_embedding");
    protos = f32c(protos_in, "prototype");
    c10::cuda::CUDAGuard guard(e.device());
    a.e = e.data_ptr<float>(), a.dim = (int32_t)e.size(1);
    if (el_in.has_value() && el_in->defined()) {
      el = f32c(*el_in, "cluster_embedding_with_loc");
      a.el = el.data_ptr<float>(), a.dim_loc = (int32_t)el.size(1);
    }
    a.protos = protos.data_ptr<float>();
    TORCH_CHECK(e.size(0) == a.n && protos.size(0) == a.m,
                "embeddings / prototypes disagree with their label vectors");
    a.status = status.data_ptr<int32_t>();
    const size_t bytes = spml_head_workspace_bytes(&a);
    state = workspace(bytes, e);
    Tensor out = at::empty({5}, e.options());
    check(spml_head_fwd(&a, state.data_ptr(), bytes, out.data_ptr<float>(), stream_of(e)),
          "spml_head_fwd");
    return out;
  }

  std::vector<Tensor> backward(const OptTensor& g_ann, const OptTensor& g_occ,
                               const OptTensor& g_sim, const OptTensor& g_total,
                               bool need_protos) {
    c10::cuda::CUDAGuard guard(e.device());
    Tensor g[4];
    const OptTensor* in[4] = {&g_ann, &g_occ, &g_sim, &g_total};
    const float* gp[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < 4; ++i)
      if (in[i]->has_value() && (*in[i])->defined()) {
        g[i] = (*in[i])->scalar_type() == at::kFloat ? **in[i] : (*in[i])->to(at::kFloat);
        gp[i] = g[i].data_ptr<float>();
      }
    const bool sim_on_el = (a.enable & 4) && !a.img_sim_on_plain;
    Tensor de = at::empty_like(e), del, dprotos;
    if (el.defined() && sim_on_el) del = at::empty_like(el);
    if (need_protos) dprotos = at::empty_like(protos);
    check(spml_head_bwd(&a, state.data_ptr(), (size_t)state.numel(), gp[0], gp[1], gp[2], gp[3],
                        de.data_ptr<float>(), del.defined() ? del.data_ptr<float>() : nullptr,
                        dprotos.defined() ? dprotos.data_ptr<float>() : nullptr, stream_of(e)),
          "spml_head_bwd");
    return {de, del, dprotos};
  }
};


// ------------------------------------------------------------------------- autograd nodes
//
// The three stage groups as C++ autograd functions: torch.autograd.Function.apply costs
// ~20 us of Python per call and the engine re-enters the interpreter for every backward; at
// batch 1 the host is the critical path of the step (profiles/r2_step_profile_b1.txt), so the
// nodes live here and Python only sees plain functions.  ops.py keeps equivalent
// torch.autograd.Function classes over ctypes for the case that this module is not built.

using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

struct SegmentParams {
  OptTensor loc, labels, sem, inst, ignore_dev, k_per_image;
  Tensor seeds, status;
  int64_t divisor, semantic_ignore, ignore_index, num_k, iterations, batch_index_offset;
  bool has_ignore;
};

struct SegmentResult {
  int64_t rows = 0, segments = 0, bits = 0;
  Tensor ibuf;
};

struct SegmentFn : public torch::autograd::Function<SegmentFn> {
  static variable_list forward(AutogradContext* ctx, const Tensor& emb, const SegmentParams* p,
                               SegmentResult* r) {
    auto res = segment_fwd(emb, p->loc, p->labels, p->sem, p->inst, p->divisor,
                           p->semantic_ignore, p->has_ignore, p->ignore_index, p->ignore_dev,
                           p->seeds, p->k_per_image, p->num_k, p->iterations,
                           p->batch_index_offset, p->status);
    std::vector<Tensor>& outs = std::get<0>(res);
    Tensor ibuf = outs.back();
    outs.pop_back();
    Tensor fbuf = outs.back();
    outs.pop_back();
    r->rows = std::get<1>(res), r->segments = std::get<2>(res), r->bits = std::get<3>(res);
    r->ibuf = ibuf;
    const int64_t B = emb.size(0);
    const int64_t loc_ch = p->loc.has_value() && p->loc->defined() ? p->loc->size(3) : 0;
    // the backward reads e, el, the two norms and the pixel -> row map: all inside these two
    // buffers (kept whole; the addresses are re-derived from the dims)
    ctx->save_for_backward({fbuf, ibuf});
    ctx->saved_data["dims"] = std::vector<int64_t>{B, emb.size(1), loc_ch, emb.size(2), emb.size(3),
                                                   (B + 2 + 4 + 3) / 4 * 4};
    ctx->mark_non_differentiable(variable_list(outs.begin() + 2, outs.end()));
    ctx->set_materialize_grads(false);
    return outs;
  }

  static variable_list backward(AutogradContext* ctx, variable_list g) {
    if (!g[0].defined() && !g[1].defined()) return {Tensor(), Tensor(), Tensor()};
    const auto saved = ctx->get_saved_variables();
    const auto d = ctx->saved_data["dims"].toIntVector();
    return {segment_bwd(saved[0], saved[1], g[0], g[1], d[0], d[1], d[2], d[3], d[4], d[5]),
            Tensor(), Tensor()};
  }
};

std::tuple<std::vector<Tensor>, int64_t, int64_t, int64_t, Tensor> segment(
    const Tensor& emb, const OptTensor& loc, const OptTensor& labels, const OptTensor& sem,
    const OptTensor& inst, int64_t divisor, int64_t semantic_ignore, bool has_ignore,
    int64_t ignore_index, const OptTensor& ignore_dev, const Tensor& seeds,
    const OptTensor& k_per_image, int64_t num_k, int64_t iterations, int64_t batch_index_offset,
    const Tensor& status) {
  SegmentParams p{loc, labels, sem, inst, ignore_dev, k_per_image, seeds, status, divisor,
                  semantic_ignore, ignore_index, num_k, iterations, batch_index_offset, has_ignore};
  SegmentResult r;
  variable_list outs = SegmentFn::apply(emb, &p, &r);
  return {outs, r.rows, r.segments, r.bits, r.ibuf};
}

struct GatherParams {
  Tensor cid, bid, sem, inst, status;
  int64_t m;
};

struct GatherFn : public torch::autograd::Function<GatherFn> {
  static variable_list forward(AutogradContext* ctx, const Tensor& e, const Tensor& el,
                               const GatherParams* p) {
    std::vector<Tensor> o = gather_fwd(e, el, p->cid, p->bid, p->sem, p->inst, p->m, p->status);
    ctx->save_for_backward({o[5], o[6]});        // fbuf, cluster ids
    ctx->saved_data["dims"] = std::vector<int64_t>{p->m, e.size(1), el.size(1)};
    ctx->mark_non_differentiable({o[2], o[3], o[4]});
    ctx->set_materialize_grads(false);   // an unused prototype set costs nothing in the backward
    o.resize(5);
    return o;
  }

  static variable_list backward(AutogradContext* ctx, variable_list g) {
    if (!g[0].defined() && !g[1].defined()) return {Tensor(), Tensor(), Tensor()};
    const auto saved = ctx->get_saved_variables();
    const auto d = ctx->saved_data["dims"].toIntVector();
    std::vector<Tensor> r = gather_bwd(saved[0], saved[1], g[0], g[1], d[0], d[1], d[2]);
    return {r[0], r[1], Tensor()};
  }
};

std::vector<Tensor> gather(const Tensor& e, const Tensor& el, const Tensor& cid, const Tensor& bid,
                           const Tensor& sem, const Tensor& inst, int64_t m, const Tensor& status) {
  GatherParams p{cid, bid, sem, inst, status, m};
  return GatherFn::apply(e, el, &p);
}

struct HeadFn : public torch::autograd::Function<HeadFn> {
  static variable_list forward(AutogradContext* ctx, const Tensor& e, const OptTensor& el,
                               const Tensor& protos, const c10::intrusive_ptr<HeadCall>& call,
                               const Tensor* status) {
    Tensor out = call->forward(e, el, protos, *status);
    ctx->saved_data["call"] = at::IValue::make_capsule(call);
    ctx->saved_data["has_el"] = el.has_value() && el->defined();
    ctx->set_materialize_grads(false);
    variable_list outs = out.unbind(0);          // sem_ann, sem_occ, img_sim, accuracy, sum
    ctx->mark_non_differentiable({outs[3]});
    return outs;
  }

  static variable_list backward(AutogradContext* ctx, variable_list g) {
    variable_list none(5);
    if (!g[0].defined() && !g[1].defined() && !g[2].defined() && !g[4].defined()) return none;
    auto call = c10::static_intrusive_pointer_cast<HeadCall>(ctx->saved_data["call"].toCapsule());
    const bool has_el = ctx->saved_data["has_el"].toBool();
    // needs_input_grad counts the tensor inputs only: e, [el,] protos
    const bool need_e = ctx->needs_input_grad(0);
    const bool need_el = has_el && ctx->needs_input_grad(1);
    const bool need_p = ctx->needs_input_grad(has_el ? 2 : 1);
    auto opt = [](const Tensor& t) { return t.defined() ? OptTensor(t) : OptTensor(); };
    std::vector<Tensor> r = call->backward(opt(g[0]), opt(g[1]), opt(g[2]), opt(g[4]), need_p);
    none[0] = need_e ? r[0] : Tensor();
    none[1] = need_el ? r[1] : Tensor();
    none[2] = r[2];
    return none;
  }
};

std::vector<Tensor> head(const Tensor& e, const OptTensor& el, const Tensor& protos,
                         const c10::intrusive_ptr<HeadCall>& call, const Tensor& status) {
  return HeadFn::apply(e, el, protos, call, &status);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "ATen-side bookkeeping around the stage-group calls of libspml_b200.so";
  m.def("abi_version", []() { return spml_abi_version(); });
  m.def("segment_fwd", &segment_fwd);
  m.def("segment_bwd", &segment_bwd);
  m.def("gather_fwd", &gather_fwd);
  m.def("gather_bwd", &gather_bwd);
  m.def("segment", &segment);
  m.def("gather", &gather);
  m.def("head", &head);
  pybind11::class_<HeadCall, c10::intrusive_ptr<HeadCall>>(m, "HeadCall")
      .def(pybind11::init([](const Tensor& cid, const OptTensor& bid, const OptTensor& sem,
                             const OptTensor& inst, const Tensor& psem, const OptTensor& pinst,
                             const OptTensor& pbid, int64_t num_classes, int64_t enable,
                             std::vector<double> kappas, std::vector<double> weights,
                             int64_t max_groups, int64_t max_rows_per_group,
                             const OptTensor& img_tags, const OptTensor& ptags, int64_t tag_col0,
                             int64_t tag_col1, const std::vector<Tensor>& bank_protos,
                             const std::vector<Tensor>& bank_sems,
                             const std::vector<Tensor>& bank_bids,
                             const std::vector<Tensor>& bank_tags,
                             const std::vector<Tensor>& bank_locs, bool nn_tags,
                             bool img_sim_on_plain, const OptTensor& protos_loc,
                             double nn_threshold) {
        return c10::make_intrusive<HeadCall>(
            cid, bid, sem, inst, psem, pinst, pbid, num_classes, enable, kappas, weights,
            max_groups, max_rows_per_group, img_tags, ptags, tag_col0, tag_col1, bank_protos,
            bank_sems, bank_bids, bank_tags, bank_locs, nn_tags, img_sim_on_plain, protos_loc,
            nn_threshold);
      }))
      .def("forward", &HeadCall::forward)
      .def("backward", &HeadCall::backward);
}
