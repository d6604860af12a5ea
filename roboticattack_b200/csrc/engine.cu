// The attack-iteration engine: one forward + input-gradient pass of OpenVLA (DINOv2 + SigLIP ViTs -> projector ->
// Llama -> action-logit loss) from the adversarial patch to d loss / d patch, as a fixed sequence of kernel launches
// on one stream over a pre-planned activation arena (no allocation, no host synchronisation, no autograd graph).
//
// Replaces the body of the reference's inner loop -- UADA.py:134-148, UADA_ddp.py:192-206, UPA.py:134-152,
// TMA.py:133-162: RandomPatchTransform -> self.vla(...) (modeling_prismatic.py:362-415 -> timm ViTs -> HF Llama)
// -> loss head -> .backward().  Only input gradients are computed (the weights are frozen, as in UADA_ddp.py:50-51);
// the unused last block of each vision tower and the lm_head rows without a label are skipped.
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vla_b200.h"
#include "kernels.h"

#define CK(expr)                  \
  do {                            \
    if (int _rc = (expr)) return _rc; \
  } while (0)

namespace {

struct Slot {
  size_t off = 0;   // element offset (bf16) into the weight arena
  int rows = 0, cols = 0;
  int loaded = 0, needed = 1;
};

struct VitDims {
  int dim, depth, heads, mlp, npre, layerscale, used, ntok, hd;
  std::string prefix;
};

struct VitBlockW {
  const bf16 *n1w, *n1b, *qkv_w, *qkv_b, *proj_w, *proj_b, *ls1, *n2w, *n2b, *fc1_w, *fc1_b, *fc2_w, *fc2_b, *ls2;
  const bf16 *qkv_t, *proj_t, *fc1_t, *fc2_t;
};
struct VitW {
  const bf16 *pe_w, *pe_b, *pe_t, *pos, *cls, *reg;
  std::vector<VitBlockW> blk;
};
struct LlamaLayerW {
  const bf16 *n1, *n2, *qkv, *o, *gu, *down, *qkv_t, *o_t, *gu_t, *down_t;
};

struct VitActs {   // per tower
  bf16* a_col;                    // im2col rows [B*np, kpad]
  std::vector<bf16*> x;           // x[i]: residual stream entering block i; x[used] = tower output   [Mv, d]
  std::vector<bf16*> qkv, attn_o, x_mid, fc1_pre;
  std::vector<float*> mean1, rstd1, mean2, rstd2, lse;
};
struct Transients {   // scratch reused across layers; one set per concurrently running chain
  bf16 *norm = nullptr, *wide = nullptr, *wide2 = nullptr, *qkv = nullptr, *a = nullptr, *b = nullptr, *c = nullptr, *d = nullptr;
  float* delta = nullptr;
};
struct LlmActs {
  std::vector<bf16*> x;           // x[l]: residual stream entering layer l; x[layers] = final  [ML, h]
  std::vector<bf16*> qkv, attn_o, x_mid, gu;
  std::vector<float*> rstd1, rstd2, lse;
};

}  // namespace

struct vla_engine {
  vla_config cfg;
  VitDims vit[2];
  int np = 0, kpad = 0, grid = 0;
  // weights
  std::unordered_map<std::string, Slot> slots;
  size_t weight_elems = 0;
  bf16* warena = nullptr;
  bool weights_resolved = false;
  VitW vw[2];
  const bf16 *pj_w[3], *pj_b[3], *pj_t[3];
  const bf16 *embed, *final_norm, *lm_head, *lm_head_t;
  std::vector<LlamaLayerW> lw;
  // plan
  int B = 0, T = 0, L = 0;
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  // batch state
  int R = 0;
  bool batch_set = false;
  int n_place = 0;
  // staging / activations
  uint8_t* obs = nullptr;
  int64_t* ids = nullptr;
  int *kv_len = nullptr, *meta = nullptr, *sup_rows = nullptr;
  int* xy = nullptr;
  float* theta = nullptr;
  float *rope_cos = nullptr, *rope_sin = nullptr;
  bool rope_set = false;
  bf16 *px = nullptr, *dpx = nullptr;
  VitActs va[2];
  bf16 *feats = nullptr, *p1_pre = nullptr, *p2_pre = nullptr;
  LlmActs la;
  bf16 *hs = nullptr, *hn = nullptr, *dlogits = nullptr;
  float *rstd_f = nullptr, *logits = nullptr, *row_stats = nullptr;
  // transients: tr[0] for the main chain (LLM, projector, DINOv2 tower), tr[1] for the SigLIP tower, which runs
  // concurrently with the DINOv2 tower on a second stream (the two towers are independent until the feature concat)
  Transients tr[2];
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool single_stream = false;
  bool fuse_swiglu_bwd = true;
  // Last decoder layer on the supervised rows only (see vla_fwd_bwd): only those rows of its output reach the loss, so
  // everything after its attention (o_proj, MLP and their backward) runs on R gathered rows instead of B*L.
  bool prune_last = true;
  // two-stream vision towers: each tower's persistent kernels take half of the SMs.  Measured SLOWER (front + ViT forward 8.2 vs
  // 7.1 ms: the towers are unbalanced, so the longer one finishes on half a machine) -> off by default, VLA_SM_SPLIT=1 enables.
  bool split_sms = false;
  bf16 *ll_xs = nullptr, *ll_as = nullptr, *ll_xm = nullptr, *ll_norm = nullptr, *ll_gu = nullptr, *ll_act = nullptr,
       *ll_dgu = nullptr, *ll_dnorm = nullptr, *ll_dxm = nullptr, *ll_dattn = nullptr;
  float* ll_rstd2 = nullptr;
  FrontendNorm nrm;
  // ---- vla_attack_step: device-resident step state + the CUDA graphs of whole attack iterations ----
  StepState* dstate = nullptr;      // device
  int* xy_cur = nullptr;            // placement of the running step (copied from xy / theta by step_begin)
  float* theta_cur = nullptr;
  float* scal_cur = nullptr;        // scalar record of the running step
  int h_place = 0, h_adam = 0;      // host mirror of dstate (advanced with every vla_attack_step)
  float h_lr = -1.f;                // learning rate currently in dstate->lr
  float* lr_ring = nullptr;         // pinned staging for learning-rate uploads
  int lr_ring_pos = 0;
  cudaStream_t cap = nullptr;       // capture stream (the caller's stream may be the legacy default stream, which cannot capture)
  struct StepGraph {
    std::vector<uint8_t> key;
    cudaGraphExec_t exec = nullptr;
    int kernel_nodes = 0, calls = 0;
  };
  std::vector<StepGraph> graphs;
  int last_kernel_nodes = 0;
  // greedy decode (vla_engine_decode_greedy)
  int* dec_ids = nullptr;           // device: the B token ids fed to the next decode step
  int* dec_state = nullptr;         // device int[2]: {cache row of the token being processed, its column in dec_tokens}
  int* dec_tokens = nullptr;        // device [B, DEC_MAX_TOKENS]: generated ids
  cudaGraphExec_t dec_graph = nullptr;   // one recorded decode step (skinny path), valid for the current plan
  int* pred_full = nullptr;         // device: argmax over the FULL vocabulary of every supervised row (what the reference's metrics use)
  std::vector<int> h_sup_rows;      // host copy of the supervised-row table of the current batch
  bool last_pass_forward_only = false;
};

namespace {

constexpr int MAX_PLACEMENTS = 4096;

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void add_slot(vla_engine* e, const std::string& name, int rows, int cols) {
  Slot s;
  s.off = e->weight_elems;
  s.rows = rows;
  s.cols = cols;
  e->slots[name] = s;
  e->weight_elems += align_up(static_cast<size_t>(rows) * cols, 128);   // 256-byte aligned slots
}

void add_matrix(vla_engine* e, const std::string& name, int out, int in) {
  add_slot(e, name, out, in);
  add_slot(e, name + "#T", in, out);
}

void build_slots(vla_engine* e) {
  const vla_config& c = e->cfg;
  for (int t = 0; t < 2; ++t) {
    const VitDims& v = e->vit[t];
    const std::string& p = v.prefix;
    add_slot(e, p + "patch_embed.proj.weight", v.dim, e->kpad);
    add_slot(e, p + "patch_embed.proj.weight#T", e->kpad, v.dim);
    add_slot(e, p + "patch_embed.proj.bias", 1, v.dim);
    add_slot(e, p + "pos_embed", e->np, v.dim);
    if (v.npre) {
      add_slot(e, p + "cls_token", 1, v.dim);
      add_slot(e, p + "reg_token", v.npre - 1, v.dim);
    }
    for (int i = 0; i < v.used; ++i) {
      const std::string b = p + "blocks." + std::to_string(i) + ".";
      add_slot(e, b + "norm1.weight", 1, v.dim);
      add_slot(e, b + "norm1.bias", 1, v.dim);
      add_matrix(e, b + "attn.qkv.weight", 3 * v.dim, v.dim);
      add_slot(e, b + "attn.qkv.bias", 1, 3 * v.dim);
      add_matrix(e, b + "attn.proj.weight", v.dim, v.dim);
      add_slot(e, b + "attn.proj.bias", 1, v.dim);
      add_slot(e, b + "norm2.weight", 1, v.dim);
      add_slot(e, b + "norm2.bias", 1, v.dim);
      add_matrix(e, b + "mlp.fc1.weight", v.mlp, v.dim);
      add_slot(e, b + "mlp.fc1.bias", 1, v.mlp);
      add_matrix(e, b + "mlp.fc2.weight", v.dim, v.mlp);
      add_slot(e, b + "mlp.fc2.bias", 1, v.dim);
      if (v.layerscale) {
        add_slot(e, b + "ls1.scale_factor", 1, v.dim);
        add_slot(e, b + "ls2.scale_factor", 1, v.dim);
      }
    }
  }
  const int vd = c.dino_dim + c.sig_dim, ph = 4 * vd, h = c.llm_hidden, f = c.llm_ffn;
  add_matrix(e, "projector.fc1.weight", ph, vd);
  add_slot(e, "projector.fc1.bias", 1, ph);
  add_matrix(e, "projector.fc2.weight", h, ph);
  add_slot(e, "projector.fc2.bias", 1, h);
  add_matrix(e, "projector.fc3.weight", h, h);
  add_slot(e, "projector.fc3.bias", 1, h);
  const std::string lm = "language_model.";
  add_slot(e, lm + "model.embed_tokens.weight", c.vocab, h);
  for (int l = 0; l < c.llm_layers; ++l) {
    const std::string p = lm + "model.layers." + std::to_string(l) + ".";
    add_slot(e, p + "input_layernorm.weight", 1, h);
    add_slot(e, p + "post_attention_layernorm.weight", 1, h);
    add_matrix(e, p + "self_attn.qkv_packed", 3 * h, h);     // filled from q_proj / k_proj / v_proj
    add_matrix(e, p + "self_attn.o_proj.weight", h, h);
    add_matrix(e, p + "mlp.gate_up_packed", 2 * f, h);       // filled from gate_proj / up_proj
    add_matrix(e, p + "mlp.down_proj.weight", h, f);
    e->slots[p + "self_attn.qkv_packed"].needed = 3;
    e->slots[p + "mlp.gate_up_packed"].needed = 2;
  }
  add_slot(e, lm + "model.norm.weight", 1, h);
  add_matrix(e, lm + "lm_head.weight", c.vocab, h);
}

const bf16* wptr(vla_engine* e, const std::string& name) {
  auto it = e->slots.find(name);
  return it == e->slots.end() ? nullptr : e->warena + it->second.off;
}

int resolve_weights(vla_engine* e) {
  for (int t = 0; t < 2; ++t) {
    const VitDims& v = e->vit[t];
    const std::string& p = v.prefix;
    VitW& w = e->vw[t];
    w.pe_w = wptr(e, p + "patch_embed.proj.weight");
    w.pe_t = wptr(e, p + "patch_embed.proj.weight#T");
    w.pe_b = wptr(e, p + "patch_embed.proj.bias");
    w.pos = wptr(e, p + "pos_embed");
    w.cls = wptr(e, p + "cls_token");
    w.reg = wptr(e, p + "reg_token");
    w.blk.resize(v.used);
    for (int i = 0; i < v.used; ++i) {
      const std::string b = p + "blocks." + std::to_string(i) + ".";
      VitBlockW& k = w.blk[i];
      k.n1w = wptr(e, b + "norm1.weight");
      k.n1b = wptr(e, b + "norm1.bias");
      k.qkv_w = wptr(e, b + "attn.qkv.weight");
      k.qkv_t = wptr(e, b + "attn.qkv.weight#T");
      k.qkv_b = wptr(e, b + "attn.qkv.bias");
      k.proj_w = wptr(e, b + "attn.proj.weight");
      k.proj_t = wptr(e, b + "attn.proj.weight#T");
      k.proj_b = wptr(e, b + "attn.proj.bias");
      k.n2w = wptr(e, b + "norm2.weight");
      k.n2b = wptr(e, b + "norm2.bias");
      k.fc1_w = wptr(e, b + "mlp.fc1.weight");
      k.fc1_t = wptr(e, b + "mlp.fc1.weight#T");
      k.fc1_b = wptr(e, b + "mlp.fc1.bias");
      k.fc2_w = wptr(e, b + "mlp.fc2.weight");
      k.fc2_t = wptr(e, b + "mlp.fc2.weight#T");
      k.fc2_b = wptr(e, b + "mlp.fc2.bias");
      k.ls1 = wptr(e, b + "ls1.scale_factor");
      k.ls2 = wptr(e, b + "ls2.scale_factor");
    }
  }
  for (int i = 0; i < 3; ++i) {
    const std::string n = "projector.fc" + std::to_string(i + 1);
    e->pj_w[i] = wptr(e, n + ".weight");
    e->pj_t[i] = wptr(e, n + ".weight#T");
    e->pj_b[i] = wptr(e, n + ".bias");
  }
  const std::string lm = "language_model.";
  e->embed = wptr(e, lm + "model.embed_tokens.weight");
  e->final_norm = wptr(e, lm + "model.norm.weight");
  e->lm_head = wptr(e, lm + "lm_head.weight");
  e->lm_head_t = wptr(e, lm + "lm_head.weight#T");
  e->lw.resize(e->cfg.llm_layers);
  for (int l = 0; l < e->cfg.llm_layers; ++l) {
    const std::string p = lm + "model.layers." + std::to_string(l) + ".";
    LlamaLayerW& w = e->lw[l];
    w.n1 = wptr(e, p + "input_layernorm.weight");
    w.n2 = wptr(e, p + "post_attention_layernorm.weight");
    w.qkv = wptr(e, p + "self_attn.qkv_packed");
    w.qkv_t = wptr(e, p + "self_attn.qkv_packed#T");
    w.o = wptr(e, p + "self_attn.o_proj.weight");
    w.o_t = wptr(e, p + "self_attn.o_proj.weight#T");
    w.gu = wptr(e, p + "mlp.gate_up_packed");
    w.gu_t = wptr(e, p + "mlp.gate_up_packed#T");
    w.down = wptr(e, p + "mlp.down_proj.weight");
    w.down_t = wptr(e, p + "mlp.down_proj.weight#T");
  }
  e->weights_resolved = true;
  return 0;
}

// ---- workspace plan ------------------------------------------------------------------------------------------
struct Bump {
  uint8_t* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

size_t plan(vla_engine* e, uint8_t* base, int B, int T) {
  const vla_config& c = e->cfg;
  Bump bp{base};
  const int H = c.img, P = e->np, L = T + P - 1;   // the last text position is never a useful query or key (see set_buffers)
  const int64_t ML = static_cast<int64_t>(B) * L;
  const int h = c.llm_hidden, f = c.llm_ffn, V = c.vocab;
  const int Rmax = B * (T - 1);
  e->obs = bp.take<uint8_t>(static_cast<size_t>(B) * H * H * 3);
  e->ids = bp.take<int64_t>(static_cast<size_t>(B) * T);
  e->kv_len = bp.take<int>(B);
  e->meta = bp.take<int>(static_cast<size_t>(Rmax) * 3);
  e->sup_rows = bp.take<int>(Rmax);
  e->xy = bp.take<int>(static_cast<size_t>(MAX_PLACEMENTS) * 2);
  e->theta = bp.take<float>(static_cast<size_t>(MAX_PLACEMENTS) * 6);
  e->rope_cos = bp.take<float>(static_cast<size_t>(L) * (h / c.llm_heads / 2));
  e->rope_sin = bp.take<float>(static_cast<size_t>(L) * (h / c.llm_heads / 2));
  e->dec_ids = bp.take<int>(B);
  e->dec_state = bp.take<int>(2);
  e->dec_tokens = bp.take<int>(static_cast<size_t>(B) * 32);
  e->pred_full = bp.take<int>(Rmax);
  e->dstate = bp.take<StepState>(1);
  e->xy_cur = bp.take<int>(static_cast<size_t>(B) * 2);
  e->theta_cur = bp.take<float>(static_cast<size_t>(B) * 6);
  e->scal_cur = bp.take<float>(LOSS_NUM_SCALARS);
  e->px = bp.take<bf16>(static_cast<size_t>(B) * 6 * H * H);
  e->dpx = bp.take<bf16>(static_cast<size_t>(B) * 6 * H * H);
  size_t max_md = 0, max_wide = 0, max_qkv = 0, max_lse = 0;
  for (int t = 0; t < 2; ++t) {
    const VitDims& v = e->vit[t];
    VitActs& a = e->va[t];
    const size_t Mv = static_cast<size_t>(B) * v.ntok;
    a.a_col = bp.take<bf16>(static_cast<size_t>(B) * P * e->kpad);
    a.x.assign(v.used + 1, nullptr);
    a.qkv.assign(v.used, nullptr);
    a.attn_o.assign(v.used, nullptr);
    a.x_mid.assign(v.used, nullptr);
    a.fc1_pre.assign(v.used, nullptr);
    a.mean1.assign(v.used, nullptr);
    a.rstd1.assign(v.used, nullptr);
    a.mean2.assign(v.used, nullptr);
    a.rstd2.assign(v.used, nullptr);
    a.lse.assign(v.used, nullptr);
    for (int i = 0; i <= v.used; ++i) a.x[i] = bp.take<bf16>(Mv * v.dim);
    for (int i = 0; i < v.used; ++i) {
      a.qkv[i] = bp.take<bf16>(Mv * 3 * v.dim);
      a.attn_o[i] = bp.take<bf16>(Mv * v.dim);
      a.x_mid[i] = bp.take<bf16>(Mv * v.dim);
      a.fc1_pre[i] = bp.take<bf16>(Mv * v.mlp);
      a.mean1[i] = bp.take<float>(Mv);
      a.rstd1[i] = bp.take<float>(Mv);
      a.mean2[i] = bp.take<float>(Mv);
      a.rstd2[i] = bp.take<float>(Mv);
      a.lse[i] = bp.take<float>(static_cast<size_t>(B) * v.heads * v.ntok);
    }
    max_md = std::max(max_md, Mv * v.dim);
    max_wide = std::max(max_wide, Mv * v.mlp);
    max_qkv = std::max(max_qkv, Mv * 3 * v.dim);
    max_lse = std::max(max_lse, static_cast<size_t>(B) * v.heads * v.ntok);
  }
  const int vd = c.dino_dim + c.sig_dim, phd = 4 * vd;
  const size_t MP = static_cast<size_t>(B) * P;
  e->feats = bp.take<bf16>(MP * vd);
  e->p1_pre = bp.take<bf16>(MP * phd);
  e->p2_pre = bp.take<bf16>(MP * h);
  max_wide = std::max(max_wide, MP * phd);
  max_md = std::max(max_md, MP * static_cast<size_t>(std::max(h, vd)));
  LlmActs& la = e->la;
  const int NL = c.llm_layers;
  la.x.assign(NL + 1, nullptr);
  la.qkv.assign(NL, nullptr);
  la.attn_o.assign(NL, nullptr);
  la.x_mid.assign(NL, nullptr);
  la.gu.assign(NL, nullptr);
  la.rstd1.assign(NL, nullptr);
  la.rstd2.assign(NL, nullptr);
  la.lse.assign(NL, nullptr);
  for (int l = 0; l <= NL; ++l) la.x[l] = bp.take<bf16>(ML * h);
  for (int l = 0; l < NL; ++l) {
    la.qkv[l] = bp.take<bf16>(ML * 3 * h);
    la.attn_o[l] = bp.take<bf16>(ML * h);
    la.x_mid[l] = bp.take<bf16>(ML * h);
    la.gu[l] = bp.take<bf16>(ML * 2 * f);
    la.rstd1[l] = bp.take<float>(ML);
    la.rstd2[l] = bp.take<float>(ML);
    la.lse[l] = bp.take<float>(static_cast<size_t>(B) * c.llm_heads * L);
  }
  max_md = std::max(max_md, static_cast<size_t>(ML) * h);
  max_wide = std::max(max_wide, static_cast<size_t>(ML) * 2 * f);
  max_qkv = std::max(max_qkv, static_cast<size_t>(ML) * 3 * h);
  max_lse = std::max(max_lse, static_cast<size_t>(B) * c.llm_heads * L);
  e->hs = bp.take<bf16>(static_cast<size_t>(Rmax) * h);
  e->hn = bp.take<bf16>(static_cast<size_t>(Rmax) * h);
  e->rstd_f = bp.take<float>(Rmax);
  e->logits = bp.take<float>(static_cast<size_t>(Rmax) * V);
  e->dlogits = bp.take<bf16>(static_cast<size_t>(Rmax) * V);
  e->row_stats = bp.take<float>(loss_head_row_stats_floats(Rmax));
  {
    const size_t Rh = static_cast<size_t>(Rmax) * h, Rf = static_cast<size_t>(Rmax) * f;
    e->ll_xs = bp.take<bf16>(Rh);
    e->ll_as = bp.take<bf16>(Rh);
    e->ll_xm = bp.take<bf16>(Rh);
    e->ll_norm = bp.take<bf16>(Rh);
    e->ll_gu = bp.take<bf16>(2 * Rf);
    e->ll_act = bp.take<bf16>(Rf);
    e->ll_dgu = bp.take<bf16>(2 * Rf);
    e->ll_dnorm = bp.take<bf16>(Rh);
    e->ll_dxm = bp.take<bf16>(Rh);
    e->ll_dattn = bp.take<bf16>(Rh);
    e->ll_rstd2 = bp.take<float>(Rmax);
  }
  {
    Transients& t0 = e->tr[0];
    t0.delta = bp.take<float>(max_lse);
    t0.norm = bp.take<bf16>(max_md);
    t0.wide = bp.take<bf16>(max_wide);
    t0.wide2 = bp.take<bf16>(max_wide);
    t0.qkv = bp.take<bf16>(max_qkv);
    t0.a = bp.take<bf16>(max_md);
    t0.b = bp.take<bf16>(max_md);
    t0.c = bp.take<bf16>(max_md);
    t0.d = bp.take<bf16>(max_md);
    const VitDims& v = e->vit[1];
    const size_t Mv = static_cast<size_t>(B) * v.ntok;
    Transients& t1 = e->tr[1];
    t1.delta = bp.take<float>(static_cast<size_t>(B) * v.heads * v.ntok);
    t1.norm = bp.take<bf16>(Mv * v.dim);
    t1.wide = bp.take<bf16>(Mv * v.mlp);
    t1.qkv = bp.take<bf16>(Mv * 3 * v.dim);
    t1.a = bp.take<bf16>(Mv * v.dim);
    t1.b = bp.take<bf16>(Mv * v.dim);
    t1.c = bp.take<bf16>(Mv * v.dim);
    t1.d = bp.take<bf16>(Mv * v.dim);
  }
  return align_up(bp.off, 256);
}

// Every W operand of the engine's GEMMs is a weight matrix (or its transpose, built at load time): a constant of the stream,
// so the kernel may request its first W tiles before griddepcontrol.wait (GemmEpilogue::w_constant).  VLA_GEMM_PREFETCH_W=0: A/B.
bool prefetch_w() {
  static const bool on = !(getenv("VLA_GEMM_PREFETCH_W") && atoi(getenv("VLA_GEMM_PREFETCH_W")) == 0);
  return on;
}
int G(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, void* out, int64_t ldc, int64_t M, int N, int K,
      const GemmEpilogue& ep, cudaStream_t s) {
  GemmEpilogue e2 = ep;
  e2.w_constant = prefetch_w();
  return gemm_bf16_tn(A, lda, W, ldw, out, ldc, static_cast<int>(M), N, K, e2, s);
}

// ---- vision tower -----------------------------------------------------------------------------------------
int vit_forward(vla_engine* e, int t, Transients& tr, cudaStream_t s) {
  const VitDims& v = e->vit[t];
  const VitW& w = e->vw[t];
  VitActs& a = e->va[t];
  const int B = e->B, d = v.dim;
  const int64_t Mv = static_cast<int64_t>(B) * v.ntok;
  CK(write_prefix_tokens(w.cls, w.reg, a.x[0], B, v.ntok, v.npre, d, s));
  {
    GemmEpilogue ep;   // conv-as-GEMM + bias, + pos_embed (broadcast over batch), rows remapped past the prefix tokens
    ep.bias = w.pe_b;
    ep.resid = w.pos;
    ep.ldr = d;
    ep.resid_mod = e->np;
    ep.out_group = e->np;
    ep.out_stride = v.ntok;
    ep.out_offset = v.npre;
    CK(G(a.a_col, e->kpad, w.pe_w, e->kpad, a.x[0], d, static_cast<int64_t>(B) * e->np, d, e->kpad, ep, s));
  }
  for (int i = 0; i < v.used; ++i) {
    const VitBlockW& k = w.blk[i];
    CK(layernorm_fwd(a.x[i], k.n1w, k.n1b, tr.norm, a.mean1[i], a.rstd1[i], Mv, d, e->cfg.vit_ln_eps, s));
    {
      GemmEpilogue ep;
      ep.bias = k.qkv_b;
      CK(G(tr.norm, d, k.qkv_w, d, a.qkv[i], 3 * d, Mv, 3 * d, d, ep, s));
    }
    CK(attention_fwd(a.qkv[i], a.attn_o[i], a.lse[i], nullptr, B, v.ntok, v.heads, v.hd, 0, s));
    {
      GemmEpilogue ep;
      ep.bias = k.proj_b;
      ep.gamma = v.layerscale ? k.ls1 : nullptr;
      ep.resid = a.x[i];
      ep.ldr = d;
      CK(G(a.attn_o[i], d, k.proj_w, d, a.x_mid[i], d, Mv, d, d, ep, s));
    }
    CK(layernorm_fwd(a.x_mid[i], k.n2w, k.n2b, tr.norm, a.mean2[i], a.rstd2[i], Mv, d, e->cfg.vit_ln_eps, s));
    {
      GemmEpilogue ep;
      ep.bias = k.fc1_b;
      ep.act = 1;
      ep.preact_out = a.fc1_pre[i];
      CK(G(tr.norm, d, k.fc1_w, d, tr.wide, v.mlp, Mv, v.mlp, d, ep, s));
    }
    {
      GemmEpilogue ep;
      ep.bias = k.fc2_b;
      ep.gamma = v.layerscale ? k.ls2 : nullptr;
      ep.resid = a.x_mid[i];
      ep.ldr = d;
      CK(G(tr.wide, v.mlp, k.fc2_w, v.mlp, a.x[i + 1], d, Mv, d, v.mlp, ep, s));
    }
  }
  return 0;
}

// dx_in: gradient wrt the tower output in tr.a ([Mv, d], prefix rows zero). Leaves d(im2col rows) in a.a_col.
int vit_backward(vla_engine* e, int t, Transients& tr, cudaStream_t s) {
  const VitDims& v = e->vit[t];
  const VitW& w = e->vw[t];
  VitActs& a = e->va[t];
  const int B = e->B, d = v.dim;
  const int64_t Mv = static_cast<int64_t>(B) * v.ntok;
  bf16* dx = tr.a;       // gradient wrt x[i+1]
  bf16* dxm = tr.b;      // gradient wrt x_mid[i]
  GemmEpilogue plain;
  // LayerScale backward (dx o gamma, the A operand of the branch's first backward GEMM) is emitted by the LayerNorm backward
  // that produces dx; only the very first one needs its own launch.
  if (v.layerscale) CK(scale_cols(dx, w.blk[v.used - 1].ls2, tr.c, Mv, d, s));
  for (int i = v.used - 1; i >= 0; --i) {
    const VitBlockW& k = w.blk[i];
    const bf16* g = v.layerscale ? tr.c : dx;
    if (v.mlp % 8 == 0) {
      GemmEpilogue ep;   // GELU backward fused into the fc2^T GEMM epilogue
      ep.aux_mode = 1;
      ep.aux = a.fc1_pre[i];
      ep.ldaux = v.mlp;
      CK(G(g, d, k.fc2_t, d, tr.wide, v.mlp, Mv, v.mlp, d, ep, s));
    } else {
      CK(G(g, d, k.fc2_t, d, tr.wide, v.mlp, Mv, v.mlp, d, plain, s));
      CK(gelu_bwd(tr.wide, a.fc1_pre[i], tr.wide, Mv * v.mlp, s));
    }
    CK(G(tr.wide, v.mlp, k.fc1_t, v.mlp, tr.norm, d, Mv, d, v.mlp, plain, s));
    CK(layernorm_bwd(tr.norm, a.x_mid[i], k.n2w, a.mean2[i], a.rstd2[i], dx, dxm, Mv, d, s, v.layerscale ? k.ls1 : nullptr,
                     v.layerscale ? tr.c : nullptr));
    g = v.layerscale ? tr.c : dxm;
    CK(G(g, d, k.proj_t, d, tr.d, d, Mv, d, d, plain, s));
    CK(attention_bwd(a.qkv[i], a.attn_o[i], tr.d, a.lse[i], tr.delta, tr.qkv, nullptr, B, v.ntok, v.heads, v.hd, 0, nullptr,
                     nullptr, 0, s));
    CK(G(tr.qkv, 3 * d, k.qkv_t, 3 * d, tr.norm, d, Mv, d, 3 * d, plain, s));
    const bool more = v.layerscale && i > 0;
    CK(layernorm_bwd(tr.norm, a.x[i], k.n1w, a.mean1[i], a.rstd1[i], dxm, dx, Mv, d, s, more ? w.blk[i - 1].ls2 : nullptr,
                     more ? tr.c : nullptr));
  }
  // patch-token rows of d x[0] -> d conv output [B*np, d] -> d im2col rows [B*np, kpad]
  CK(copy_rows(dx, d, v.ntok, v.npre, tr.c, d, e->np, 0, B, e->np, d, s));
  CK(G(tr.c, d, w.pe_t, d, a.a_col, e->kpad, static_cast<int64_t>(B) * e->np, e->kpad, d, plain, s));
  return 0;
}

// ---- both vision towers in lock step: the same-depth kernels of DINOv2 and SigLIP share one launch --------------------
// The towers are independent until the feature concat, but every one of their kernels is a short (10-30 us), latency-bound
// launch that fills at most 1-3 waves of the machine; two streams do not help because the persistent GEMM / attention kernels
// occupy every SM (one CTA of ~200 KB shared memory each) and therefore serialise anyway.  Issuing the DINOv2 and the SigLIP
// GEMM (and LayerNorm) of the same depth as ONE launch pays the prologue / pipeline fill / drain once and lets a tile's
// epilogue overlap the next tile's main loop.  SigLIP's three extra blocks run as single-problem launches.
struct TowerCtx {
  const VitDims* v;
  const VitW* w;
  VitActs* a;
  Transients* tr;
  int64_t Mv;
};

GemmProblem GP(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, void* out, int64_t ldc, int64_t M, int N, int K, const GemmEpilogue& ep) {
  return GemmProblem{A, lda, W, ldw, out, ldc, static_cast<int>(M), N, K, ep};
}
// runs problem builder f(tower) for the towers that still have block i: one dual launch, or one single launch
template <typename F>
int gemm_towers(const TowerCtx (&tc)[2], int i, F f, cudaStream_t s) {
  const bool h0 = i < tc[0].v->used, h1 = i < tc[1].v->used;
  if (h0 && h1) {
    GemmProblem p0 = f(tc[0]), p1 = f(tc[1]);
    p0.epi.w_constant = p1.epi.w_constant = prefetch_w();
    return gemm_bf16_tn_dual(p0, p1, s);
  }
  GemmProblem p = f(tc[h0 ? 0 : 1]);
  p.epi.w_constant = prefetch_w();
  return gemm_bf16_tn(p.A, p.lda, p.W, p.ldw, p.out, p.ldc, p.M, p.N, p.K, p.epi, s);
}

int vit_forward_both(vla_engine* e, cudaStream_t s) {
  const int B = e->B;
  TowerCtx tc[2];
  for (int t = 0; t < 2; ++t) tc[t] = TowerCtx{&e->vit[t], &e->vw[t], &e->va[t], &e->tr[t], static_cast<int64_t>(B) * e->vit[t].ntok};
  const float eps = e->cfg.vit_ln_eps;
  for (int t = 0; t < 2; ++t)
    CK(write_prefix_tokens(tc[t].w->cls, tc[t].w->reg, tc[t].a->x[0], B, tc[t].v->ntok, tc[t].v->npre, tc[t].v->dim, s));
  CK(gemm_towers(tc, 0, [&](const TowerCtx& c) {
    GemmEpilogue ep;   // conv-as-GEMM + bias, + pos_embed (broadcast over batch), rows remapped past the prefix tokens
    ep.bias = c.w->pe_b;
    ep.resid = c.w->pos;
    ep.ldr = c.v->dim;
    ep.resid_mod = e->np;
    ep.out_group = e->np;
    ep.out_stride = c.v->ntok;
    ep.out_offset = c.v->npre;
    return GP(c.a->a_col, e->kpad, c.w->pe_w, e->kpad, c.a->x[0], c.v->dim, static_cast<int64_t>(B) * e->np, c.v->dim, e->kpad, ep);
  }, s));
  const int nb = std::max(tc[0].v->used, tc[1].v->used);
  auto ln_fwd = [&](int i, bool second) -> int {
    LnFwdProblem p[2];
    int n = 0;
    for (int t = 0; t < 2; ++t) {
      const TowerCtx& c = tc[t];
      if (i >= c.v->used) continue;
      const VitBlockW& k = c.w->blk[i];
      p[n++] = second ? LnFwdProblem{c.a->x_mid[i], k.n2w, k.n2b, c.tr->norm, c.a->mean2[i], c.a->rstd2[i], c.Mv, c.v->dim}
                      : LnFwdProblem{c.a->x[i], k.n1w, k.n1b, c.tr->norm, c.a->mean1[i], c.a->rstd1[i], c.Mv, c.v->dim};
    }
    if (n == 2) return layernorm_fwd2(p[0], p[1], eps, s);
    return layernorm_fwd(p[0].x, p[0].w, p[0].b, p[0].y, p[0].mean, p[0].rstd, p[0].M, p[0].d, eps, s);
  };
  for (int i = 0; i < nb; ++i) {
    CK(ln_fwd(i, false));
    CK(gemm_towers(tc, i, [&](const TowerCtx& c) {
      const int d = c.v->dim;
      GemmEpilogue ep;
      ep.bias = c.w->blk[i].qkv_b;
      return GP(c.tr->norm, d, c.w->blk[i].qkv_w, d, c.a->qkv[i], 3 * d, c.Mv, 3 * d, d, ep);
    }, s));
    for (int t = 0; t < 2; ++t)
      if (i < tc[t].v->used)
        CK(attention_fwd(tc[t].a->qkv[i], tc[t].a->attn_o[i], tc[t].a->lse[i], nullptr, B, tc[t].v->ntok, tc[t].v->heads, tc[t].v->hd, 0, s));
    CK(gemm_towers(tc, i, [&](const TowerCtx& c) {
      const int d = c.v->dim;
      GemmEpilogue ep;
      ep.bias = c.w->blk[i].proj_b;
      ep.gamma = c.v->layerscale ? c.w->blk[i].ls1 : nullptr;
      ep.resid = c.a->x[i];
      ep.ldr = d;
      return GP(c.a->attn_o[i], d, c.w->blk[i].proj_w, d, c.a->x_mid[i], d, c.Mv, d, d, ep);
    }, s));
    CK(ln_fwd(i, true));
    CK(gemm_towers(tc, i, [&](const TowerCtx& c) {
      const int d = c.v->dim;
      GemmEpilogue ep;
      ep.bias = c.w->blk[i].fc1_b;
      ep.act = 1;
      ep.preact_out = c.a->fc1_pre[i];
      return GP(c.tr->norm, d, c.w->blk[i].fc1_w, d, c.tr->wide, c.v->mlp, c.Mv, c.v->mlp, d, ep);
    }, s));
    CK(gemm_towers(tc, i, [&](const TowerCtx& c) {
      const int d = c.v->dim;
      GemmEpilogue ep;
      ep.bias = c.w->blk[i].fc2_b;
      ep.gamma = c.v->layerscale ? c.w->blk[i].ls2 : nullptr;
      ep.resid = c.a->x_mid[i];
      ep.ldr = d;
      return GP(c.tr->wide, c.v->mlp, c.w->blk[i].fc2_w, c.v->mlp, c.a->x[i + 1], d, c.Mv, d, c.v->mlp, ep);
    }, s));
  }
  return 0;
}

// Gradients wrt the tower outputs in tr[t].a ([Mv, d], prefix rows zero).  Leaves d(im2col rows) in va[t].a_col.
int vit_backward_both(vla_engine* e, cudaStream_t s) {
  const int B = e->B;
  TowerCtx tc[2];
  for (int t = 0; t < 2; ++t) tc[t] = TowerCtx{&e->vit[t], &e->vw[t], &e->va[t], &e->tr[t], static_cast<int64_t>(B) * e->vit[t].ntok};
  GemmEpilogue plain;
  // per tower: dx = tr.a (gradient wrt x[i+1]), dxm = tr.b (wrt x_mid[i]); LayerScale towers keep dx o gamma in tr.c
  for (int t = 0; t < 2; ++t)
    if (tc[t].v->layerscale)
      CK(scale_cols(tc[t].tr->a, tc[t].w->blk[tc[t].v->used - 1].ls2, tc[t].tr->c, tc[t].Mv, tc[t].v->dim, s));
  // all towers' MLP widths must allow the fused GELU backward for the dual launch to share one epilogue kind
  const bool fuse_gelu = tc[0].v->mlp % 8 == 0 && tc[1].v->mlp % 8 == 0;
  const int nb = std::max(tc[0].v->used, tc[1].v->used);
  auto ln_bwd = [&](int i, bool first) -> int {   // first: norm1 (end of the block's backward), else norm2
    LnBwdProblem p[2];
    int n = 0;
    for (int t = 0; t < 2; ++t) {
      const TowerCtx& c = tc[t];
      if (i >= c.v->used) continue;
      const VitBlockW& k = c.w->blk[i];
      bf16 *dx = c.tr->a, *dxm = c.tr->b;
      if (!first) {
        p[n++] = LnBwdProblem{c.tr->norm, c.a->x_mid[i], k.n2w, c.a->mean2[i], c.a->rstd2[i], dx, dxm, c.Mv, c.v->dim,
                              c.v->layerscale ? k.ls1 : nullptr, c.v->layerscale ? c.tr->c : nullptr};
      } else {
        const bool more = c.v->layerscale && i > 0;
        p[n++] = LnBwdProblem{c.tr->norm, c.a->x[i], k.n1w, c.a->mean1[i], c.a->rstd1[i], dxm, dx, c.Mv, c.v->dim,
                              more ? c.w->blk[i - 1].ls2 : nullptr, more ? c.tr->c : nullptr};
      }
    }
    if (n == 2) return layernorm_bwd2(p[0], p[1], s);
    return layernorm_bwd(p[0].dy, p[0].x, p[0].w, p[0].mean, p[0].rstd, p[0].dres, p[0].dx, p[0].M, p[0].d, s, p[0].gamma2, p[0].scaled);
  };
  for (int i = nb - 1; i >= 0; --i) {
    CK(gemm_towers(tc, i, [&](const TowerCtx& c) {   // d(fc2 input) = g . W_fc2, GELU backward fused
      const int d = c.v->dim;
      const bf16* g = c.v->layerscale ? c.tr->c : c.tr->a;
      GemmEpilogue ep;
      if (fuse_gelu) {
        ep.aux_mode = 1;
        ep.aux = c.a->fc1_pre[i];
        ep.ldaux = c.v->mlp;
      }
      return GP(g, d, c.w->blk[i].fc2_t, d, c.tr->wide, c.v->mlp, c.Mv, c.v->mlp, d, ep);
    }, s));
    if (!fuse_gelu)
      for (int t = 0; t < 2; ++t)
        if (i < tc[t].v->used) CK(gelu_bwd(tc[t].tr->wide, tc[t].a->fc1_pre[i], tc[t].tr->wide, tc[t].Mv * tc[t].v->mlp, s));
    CK(gemm_towers(tc, i, [&](const TowerCtx& c) {
      return GP(c.tr->wide, c.v->mlp, c.w->blk[i].fc1_t, c.v->mlp, c.tr->norm, c.v->dim, c.Mv, c.v->dim, c.v->mlp, plain);
    }, s));
    CK(ln_bwd(i, false));
    CK(gemm_towers(tc, i, [&](const TowerCtx& c) {
      const int d = c.v->dim;
      const bf16* g = c.v->layerscale ? c.tr->c : c.tr->b;
      return GP(g, d, c.w->blk[i].proj_t, d, c.tr->d, d, c.Mv, d, d, plain);
    }, s));
    for (int t = 0; t < 2; ++t)
      if (i < tc[t].v->used)
        CK(attention_bwd(tc[t].a->qkv[i], tc[t].a->attn_o[i], tc[t].tr->d, tc[t].a->lse[i], tc[t].tr->delta, tc[t].tr->qkv, nullptr, B,
                         tc[t].v->ntok, tc[t].v->heads, tc[t].v->hd, 0, nullptr, nullptr, 0, s));
    CK(gemm_towers(tc, i, [&](const TowerCtx& c) {
      const int d = c.v->dim;
      return GP(c.tr->qkv, 3 * d, c.w->blk[i].qkv_t, 3 * d, c.tr->norm, d, c.Mv, d, 3 * d, plain);
    }, s));
    CK(ln_bwd(i, true));
  }
  // patch-token rows of d x[0] -> d conv output [B*np, d] -> d im2col rows [B*np, kpad]
  for (int t = 0; t < 2; ++t)
    CK(copy_rows(tc[t].tr->a, tc[t].v->dim, tc[t].v->ntok, tc[t].v->npre, tc[t].tr->c, tc[t].v->dim, e->np, 0, B, e->np, tc[t].v->dim, s));
  CK(gemm_towers(tc, 0, [&](const TowerCtx& c) {
    return GP(c.tr->c, c.v->dim, c.w->pe_t, c.v->dim, c.a->a_col, e->kpad, static_cast<int64_t>(B) * e->np, e->kpad, c.v->dim, plain);
  }, s));
  return 0;
}

}  // namespace

// =================================================================================================================
// C ABI
// =================================================================================================================
extern "C" int vla_engine_create(const vla_config* cfg, vla_engine** out) {
  VLA_REQUIRE(cfg && out, "vla_engine_create: null argument");
  VLA_REQUIRE(cfg->img % cfg->patch == 0, "image size %d is not a multiple of the ViT patch %d", cfg->img, cfg->patch);
  VLA_REQUIRE(cfg->llm_hidden % cfg->llm_heads == 0 && cfg->dino_dim % cfg->dino_heads == 0 &&
                  cfg->sig_dim % cfg->sig_heads == 0,
              "hidden sizes must be divisible by head counts");
  VLA_REQUIRE(cfg->vocab >= 32000, "vocabulary must contain the action ids 31744..31999");
  vla_engine* e = new vla_engine();
  e->cfg = *cfg;
  e->grid = cfg->img / cfg->patch;
  e->np = e->grid * e->grid;
  e->kpad = static_cast<int>(align_up(3 * cfg->patch * cfg->patch, 64));
  e->vit[0] = {cfg->dino_dim, cfg->dino_depth, cfg->dino_heads, cfg->dino_mlp, cfg->dino_prefix, cfg->dino_layerscale,
               cfg->dino_depth - 1, e->np + cfg->dino_prefix, cfg->dino_dim / cfg->dino_heads,
               "vision_backbone.featurizer."};
  e->vit[1] = {cfg->sig_dim, cfg->sig_depth, cfg->sig_heads, cfg->sig_mlp, cfg->sig_prefix, cfg->sig_layerscale,
               cfg->sig_depth - 1, e->np + cfg->sig_prefix, cfg->sig_dim / cfg->sig_heads,
               "vision_backbone.fused_featurizer."};
  for (int s = 0; s < 2; ++s)
    for (int c = 0; c < 3; ++c) {
      e->nrm.mean[s][c] = cfg->norm_mean[s][c];
      e->nrm.std[s][c] = cfg->norm_std[s][c];
    }
  build_slots(e);
  *out = e;
  return 0;
}

extern "C" void vla_engine_destroy(vla_engine* e) {
  if (!e) return;
  if (e->side) {
    cudaStreamDestroy(e->side);
    cudaEventDestroy(e->ev_fork);
    cudaEventDestroy(e->ev_join);
  }
  for (auto& g : e->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (e->dec_graph) cudaGraphExecDestroy(e->dec_graph);
  if (e->cap) cudaStreamDestroy(e->cap);
  if (e->lr_ring) cudaFreeHost(e->lr_ring);
  delete e;
}

extern "C" size_t vla_engine_weight_bytes(const vla_engine* e) { return e->weight_elems * sizeof(bf16); }

extern "C" size_t vla_engine_workspace_bytes(vla_engine* e, int B, int T) {
  vla_engine tmp = *e;   // plan() on a copy with a null base only measures
  return plan(&tmp, nullptr, B, T);
}

extern "C" int vla_engine_set_buffers(vla_engine* e, void* weight_arena, size_t weight_bytes, void* workspace,
                                      size_t workspace_bytes, int B, int T) {
  VLA_REQUIRE(e && weight_arena && workspace, "vla_engine_set_buffers: null argument");
  VLA_REQUIRE(B >= 1 && T >= 2, "vla_engine_set_buffers: need B >= 1 and T >= 2 (got %d, %d)", B, T);
  VLA_REQUIRE(weight_bytes >= vla_engine_weight_bytes(e), "weight arena too small: %zu < %zu", weight_bytes,
              vla_engine_weight_bytes(e));
  const size_t need = vla_engine_workspace_bytes(e, B, T);
  VLA_REQUIRE(workspace_bytes >= need, "workspace too small: %zu < %zu", workspace_bytes, need);
  VLA_REQUIRE((reinterpret_cast<uintptr_t>(weight_arena) & 255) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "arenas must be 256-byte aligned");
  e->warena = static_cast<bf16*>(weight_arena);
  e->ws = static_cast<uint8_t*>(workspace);
  e->ws_bytes = workspace_bytes;
  e->B = B;
  e->T = T;
  // Sequence run through the LLM: [BOS | np patch rows | text 1..T-2].  The last text position T-1 is dropped: under
  // the causal mask no other position attends to it and the logits row it would produce predicts nothing (the shifted
  // CE pairs logits[:, :-1] with labels[:, 1:]), so the loss and every gradient are unchanged.
  e->L = T + e->np - 1;
  plan(e, e->ws, B, T);
  if (!e->side) {
    VLA_CHECK_CUDA(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
    VLA_CHECK_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    VLA_CHECK_CUDA(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    const char* ss = getenv("VLA_SINGLE_STREAM");
    e->single_stream = ss && atoi(ss) != 0;
    const char* fs = getenv("VLA_FUSE_SWIGLU_BWD");
    e->fuse_swiglu_bwd = !(fs && atoi(fs) == 0);
    const char* sp = getenv("VLA_SM_SPLIT");
    e->split_sms = sp && atoi(sp) != 0;
    const char* pl = getenv("VLA_PRUNE_LAST");
    e->prune_last = !(pl && atoi(pl) == 0);
  }
  e->batch_set = false;
  e->rope_set = false;
  e->n_place = 0;
  for (auto& g : e->graphs)   // the plan moved: recorded graphs hold stale pointers / shapes
    if (g.exec) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
  if (e->dec_graph) {
    cudaGraphExecDestroy(e->dec_graph);
    e->dec_graph = nullptr;
  }
  e->h_place = e->h_adam = 0;
  e->h_lr = -1.f;
  VLA_CHECK_CUDA(cudaMemset(e->dstate, 0, sizeof(StepState)));
  VLA_CHECK_CUDA(cudaMemset(e->scal_cur, 0, sizeof(float) * LOSS_NUM_SCALARS));
  return resolve_weights(e);
}

// Copies one checkpoint tensor (bf16, device memory, HF name) into the arena, building the derived forms the kernels
// use: K-padded conv weight, packed q|k|v and gate|up, and the transposed copy every input-gradient GEMM reads.
extern "C" int vla_engine_load_weight(vla_engine* e, const char* name_c, const void* src_v, int64_t numel, void* stream) {
  VLA_REQUIRE(e && e->warena, "vla_engine_load_weight: call vla_engine_set_buffers first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bf16* src = static_cast<const bf16*>(src_v);
  std::string name(name_c);
  const int h = e->cfg.llm_hidden, f = e->cfg.llm_ffn;
  auto find = [&](const std::string& n) -> Slot* {
    auto it = e->slots.find(n);
    return it == e->slots.end() ? nullptr : &it->second;
  };
  // --- packed Llama projections ---
  struct Pack { const char* suffix; const char* packed; int index; int rows; };
  const Pack packs[] = {{"self_attn.q_proj.weight", "self_attn.qkv_packed", 0, h},
                        {"self_attn.k_proj.weight", "self_attn.qkv_packed", 1, h},
                        {"self_attn.v_proj.weight", "self_attn.qkv_packed", 2, h},
                        {"mlp.gate_proj.weight", "mlp.gate_up_packed", 0, f},
                        {"mlp.up_proj.weight", "mlp.gate_up_packed", 1, f}};
  for (const Pack& pk : packs) {
    const size_t sl = strlen(pk.suffix);
    if (name.size() > sl && name.compare(name.size() - sl, sl, pk.suffix) == 0) {
      const std::string pname = name.substr(0, name.size() - sl) + pk.packed;
      Slot* sp = find(pname);
      Slot* st = find(pname + "#T");
      VLA_REQUIRE(sp && st, "unknown weight '%s'", name_c);
      VLA_REQUIRE(numel == static_cast<int64_t>(pk.rows) * h, "weight '%s': expected %lld elements, got %lld", name_c,
                  static_cast<long long>(pk.rows) * h, static_cast<long long>(numel));
      const bool gate_up = pk.rows == f && strstr(pk.suffix, "mlp.") != nullptr;
      if (!gate_up) {
        // q | k | v stacked row-wise: block `index` occupies rows [index*rows, (index+1)*rows)
        bf16* dst = e->warena + sp->off + static_cast<size_t>(pk.index) * pk.rows * h;
        VLA_CHECK_CUDA(cudaMemcpyAsync(dst, src, numel * sizeof(bf16), cudaMemcpyDeviceToDevice, s));
      } else {
        // gate / up interleaved in groups of 64 output features: packed rows [128k, 128k+64) = gate rows [64k, 64k+64),
        // [128k+64, 128k+128) = up rows -- so that one epilogue thread of the gate|up GEMM holds gate_i and up_i
        VLA_REQUIRE(f % 64 == 0, "llm_ffn must be a multiple of 64 (got %d)", f);
        bf16* dst = e->warena + sp->off + static_cast<size_t>(pk.index) * 64 * h;
        VLA_CHECK_CUDA(cudaMemcpy2DAsync(dst, static_cast<size_t>(128) * h * sizeof(bf16), src, static_cast<size_t>(64) * h * sizeof(bf16),
                                         static_cast<size_t>(64) * h * sizeof(bf16), f / 64, cudaMemcpyDeviceToDevice, s));
      }
      sp->loaded++;
      st->loaded = sp->loaded;
      if (sp->loaded == sp->needed)   // all parts in: one transpose of the packed matrix [rows, h] -> [h, rows]
        CK(transpose_bf16(e->warena + sp->off, sp->cols, e->warena + st->off, st->cols, sp->rows, sp->cols, s));
      return 0;
    }
  }
  Slot* sp = find(name);
  VLA_REQUIRE(sp, "unknown weight '%s' (not on the attack hot path)", name_c);
  bf16* dst = e->warena + sp->off;
  const std::string pe = "patch_embed.proj.weight";
  if (name.size() > pe.size() && name.compare(name.size() - pe.size(), pe.size(), pe) == 0) {
    const int kk = 3 * e->cfg.patch * e->cfg.patch;
    VLA_REQUIRE(numel == static_cast<int64_t>(sp->rows) * kk, "weight '%s': expected %lld elements, got %lld", name_c,
                static_cast<long long>(sp->rows) * kk, static_cast<long long>(numel));
    VLA_CHECK_CUDA(cudaMemsetAsync(dst, 0, static_cast<size_t>(sp->rows) * sp->cols * sizeof(bf16), s));
    VLA_CHECK_CUDA(cudaMemcpy2DAsync(dst, sp->cols * sizeof(bf16), src, kk * sizeof(bf16), kk * sizeof(bf16), sp->rows,
                                     cudaMemcpyDeviceToDevice, s));
  } else {
    VLA_REQUIRE(numel == static_cast<int64_t>(sp->rows) * sp->cols, "weight '%s': expected %lld elements, got %lld", name_c,
                static_cast<long long>(sp->rows) * sp->cols, static_cast<long long>(numel));
    VLA_CHECK_CUDA(cudaMemcpyAsync(dst, src, numel * sizeof(bf16), cudaMemcpyDeviceToDevice, s));
  }
  sp->loaded++;
  if (Slot* st = find(name + "#T")) {
    CK(transpose_bf16(dst, sp->cols, e->warena + st->off, st->cols, sp->rows, sp->cols, s));
    st->loaded = sp->loaded;
  }
  return 0;
}

extern "C" int vla_engine_weights_ready(vla_engine* e) {
  int missing = 0;
  std::string first;
  for (auto& kv : e->slots)
    if (kv.second.loaded < kv.second.needed) {
      if (!missing) first = kv.first;
      ++missing;
    }
  VLA_REQUIRE(missing == 0, "%d weight tensors not loaded (e.g. '%s')", missing, first.c_str());
  return 0;
}

// cos / sin tables [L, head_dim/2] fp32 (host), computed by the caller exactly as HF LlamaRotaryEmbedding does
// (fp32 outer product, cos/sin, cast to bf16) so that the engine and the oracle share them bit for bit.
extern "C" int vla_engine_set_rope(vla_engine* e, const float* cos_host, const float* sin_host, int L, void* stream) {
  VLA_REQUIRE(e && e->ws, "vla_engine_set_rope: call vla_engine_set_buffers first");
  VLA_REQUIRE(L == e->L, "rope table length %d != planned sequence length %d", L, e->L);
  const size_t n = static_cast<size_t>(L) * (e->cfg.llm_hidden / e->cfg.llm_heads / 2) * sizeof(float);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VLA_CHECK_CUDA(cudaMemcpyAsync(e->rope_cos, cos_host, n, cudaMemcpyHostToDevice, s));
  VLA_CHECK_CUDA(cudaMemcpyAsync(e->rope_sin, sin_host, n, cudaMemcpyHostToDevice, s));
  e->rope_set = true;
  return 0;
}

// One outer iteration's batch (UADA.py:120-130): host tensors as the collator yields them (data_utils.py:183-217).
// The clean observations stay on the device for all inner steps; supervised rows are located on the host once.
extern "C" int vla_engine_set_batch(vla_engine* e, const uint8_t* obs, int obs_on_device, const int64_t* input_ids,
                                    const uint8_t* attention_mask, const int64_t* labels, int B, int T, void* stream) {
  VLA_REQUIRE(e && e->ws, "vla_engine_set_batch: call vla_engine_set_buffers first");
  VLA_REQUIRE(B == e->B && T == e->T, "batch shape (%d, %d) != planned (%d, %d)", B, T, e->B, e->T);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int P = e->np, L = e->L;
  std::vector<int> kv(B), meta, rows;
  for (int b = 0; b < B; ++b) {
    int n = 0;
    bool ended = false;
    for (int t = 0; t < T; ++t) {
      if (attention_mask[b * T + t]) {
        VLA_REQUIRE(!ended, "attention_mask of sample %d is not a right-padded prefix", b);
        ++n;
      } else {
        ended = true;
      }
      const int64_t id = input_ids[b * T + t];
      VLA_REQUIRE(id >= 0 && id < e->cfg.vocab, "input id %lld out of range", static_cast<long long>(id));
    }
    VLA_REQUIRE(n >= 1, "sample %d has an empty attention mask", b);
    kv[b] = std::min(P + n, L);
    int idx = 0;
    for (int t = 1; t < T; ++t) {
      const int64_t y = labels[b * T + t];
      if (y == -100) continue;
      VLA_REQUIRE(y >= 0 && y < e->cfg.vocab, "label %lld out of range", static_cast<long long>(y));
      rows.push_back(b * L + P + t - 1);   // logits row that predicts text position t (shifted CE)
      meta.push_back(static_cast<int>(y));
      meta.push_back(b);
      meta.push_back(idx++);
    }
  }
  e->R = static_cast<int>(rows.size());
  VLA_REQUIRE(e->R > 0, "no supervised tokens in the batch (all labels are -100)");
  const size_t obs_bytes = static_cast<size_t>(B) * e->cfg.img * e->cfg.img * 3;
  VLA_CHECK_CUDA(cudaMemcpyAsync(e->obs, obs, obs_bytes, obs_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  VLA_CHECK_CUDA(cudaMemcpyAsync(e->ids, input_ids, sizeof(int64_t) * B * T, cudaMemcpyHostToDevice, s));
  VLA_CHECK_CUDA(cudaMemcpyAsync(e->kv_len, kv.data(), sizeof(int) * B, cudaMemcpyHostToDevice, s));
  VLA_CHECK_CUDA(cudaMemcpyAsync(e->meta, meta.data(), sizeof(int) * meta.size(), cudaMemcpyHostToDevice, s));
  VLA_CHECK_CUDA(cudaMemcpyAsync(e->sup_rows, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice, s));
  VLA_CHECK_CUDA(cudaStreamSynchronize(s));   // the std::vectors above are pageable staging
  e->h_sup_rows = rows;
  e->last_pass_forward_only = false;
  e->batch_set = true;
  return 0;
}

// Placements for the next `nsteps` inner iterations, drawn by the host RNG protocol (appply_random_transform.py:123-129)
// in one go: xy int32 [nsteps, B, 2], theta float32 [nsteps, B, 2, 3].
extern "C" int vla_engine_set_placements(vla_engine* e, const int* xy_host, const float* theta_host, int nsteps, void* stream) {
  VLA_REQUIRE(e && e->ws, "vla_engine_set_placements: call vla_engine_set_buffers first");
  VLA_REQUIRE(nsteps >= 1 && nsteps * e->B <= MAX_PLACEMENTS, "too many placements: %d steps x %d images > %d", nsteps, e->B,
              MAX_PLACEMENTS);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VLA_CHECK_CUDA(cudaMemcpyAsync(e->xy, xy_host, sizeof(int) * 2 * nsteps * e->B, cudaMemcpyHostToDevice, s));
  VLA_CHECK_CUDA(cudaMemcpyAsync(e->theta, theta_host, sizeof(float) * 6 * nsteps * e->B, cudaMemcpyHostToDevice, s));
  VLA_CHECK_CUDA(cudaStreamSynchronize(s));
  e->n_place = nsteps;
  return 0;
}

extern "C" int vla_engine_num_supervised(const vla_engine* e) { return e->R; }

// ids i32 [num_supervised] (device): argmax over the full vocabulary of every supervised logits row of the last pass
extern "C" int vla_engine_full_vocab_pred(vla_engine* e, int* dst, void* stream) {
  VLA_REQUIRE(e && e->ws && dst, "vla_engine_full_vocab_pred: null argument");
  VLA_CHECK_CUDA(cudaMemcpyAsync(dst, e->pred_full, sizeof(int) * e->R, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return 0;
}

// 1 = keep both vision towers on the caller's stream (clean per-kernel timing for profiling); 0 = fork the SigLIP
// tower onto the engine's side stream (default).
extern "C" int vla_engine_set_single_stream(vla_engine* e, int on) {
  VLA_REQUIRE(e != nullptr, "vla_engine_set_single_stream: null engine");
  e->single_stream = on != 0;
  return 0;
}

namespace {
bool phase_timing_on() {
  static const bool on = getenv("VLA_PHASE_TIMING") && atoi(getenv("VLA_PHASE_TIMING")) != 0;
  return on;
}
int fwd_bwd_impl(vla_engine* e, const float* patch, int ph, int pw, const int* xy, const float* th, int fe_mode,
                 const vla_loss_params* lp_c, float* dpatch, float* scalars, int* pred_ids, int flags, cudaStream_t s);
}  // namespace

extern "C" int vla_fwd_bwd(vla_engine* e, const float* patch, int ph, int pw, int step_idx, int fe_mode,
                           const vla_loss_params* lp_c, float* dpatch, float* scalars, int* pred_ids, int flags,
                           void* stream) {
  VLA_REQUIRE(e && e->weights_resolved, "vla_fwd_bwd: engine not initialised");
  VLA_REQUIRE(e->batch_set, "vla_fwd_bwd: call vla_engine_set_batch first");
  VLA_REQUIRE(e->rope_set, "vla_fwd_bwd: call vla_engine_set_rope first");
  VLA_REQUIRE(patch && lp_c && dpatch && scalars && pred_ids, "vla_fwd_bwd: null argument");
  VLA_REQUIRE(fe_mode == FE_MODE_NONE || (step_idx >= 0 && step_idx < e->n_place),
              "vla_fwd_bwd: placement %d not uploaded (have %d)", step_idx, e->n_place);
  const int* xy = e->xy + static_cast<size_t>(step_idx < 0 ? 0 : step_idx) * e->B * 2;
  const float* th = e->theta + static_cast<size_t>(step_idx < 0 ? 0 : step_idx) * e->B * 6;
  return fwd_bwd_impl(e, patch, ph, pw, xy, th, fe_mode, lp_c, dpatch, scalars, pred_ids, flags, static_cast<cudaStream_t>(stream));
}

namespace {
int fwd_bwd_impl(vla_engine* e, const float* patch, int ph, int pw, const int* xy, const float* th, int fe_mode,
                 const vla_loss_params* lp_c, float* dpatch, float* scalars, int* pred_ids, int flags, cudaStream_t s) {
  const vla_config& c = e->cfg;
  const int B = e->B, T = e->T, L = e->L, P = e->np, H = c.img;
  const int h = c.llm_hidden, f = c.llm_ffn, V = c.vocab, NH = c.llm_heads, hd = h / NH;
  const int64_t ML = static_cast<int64_t>(B) * L;
  const int64_t MP = static_cast<int64_t>(B) * P;
  const int vd = c.dino_dim + c.sig_dim, phd = 4 * vd;
  LossParams lp{lp_c->kind, lp_c->mse_weight, lp_c->alpha, lp_c->belta, lp_c->ce_scale};
  GemmEpilogue plain;

  // VLA_PHASE_TIMING=1 (diagnostics): events at the phase boundaries, printed after a synchronise
  const bool phase_timing = phase_timing_on();
  static cudaEvent_t pev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  auto mark = [&](int i) {
    if (!phase_timing) return;
    if (!pev[i]) cudaEventCreate(&pev[i]);
    cudaEventRecord(pev[i], s);
  };
  mark(0);

  // ---------------- forward ----------------
  CK(patch_frontend_fwd(e->obs, patch, xy, th, e->px, B, H, H, ph, pw, fe_mode, e->nrm, s));
  CK(im2col_patches(e->px, e->va[0].a_col, e->va[1].a_col, B, H, H, c.patch, e->kpad, s));
  // The two towers are independent until the feature concat: DINOv2 stays on the caller's stream, SigLIP runs on the
  // engine's side stream (fork / join with events), so the many small kernels of one tower fill the launch gaps, ramps
  // and tails of the other.  VLA_SINGLE_STREAM=1 keeps everything on one stream.
  // Default: both towers' kernels of the same depth in one launch (lock step) -- except for a single sample, where a tower's
  // GEMMs have a handful of tiles and the two towers on two streams fill more of the machine (18.17 -> 17.63 ms per iteration
  // at bs 1; 25.20 / 25.11 at bs 2, 37.20 / 37.05 at bs 4, 66.2 / 66.5 at bs 8: no difference from two samples on).
  const char* towers_env = getenv("VLA_TOWERS");   // read per call (A/B switch, tests): "streams" / "lockstep" force one form
  const bool towers_dual = towers_env ? strcmp(towers_env, "streams") != 0 : B > 1;
  const bool two_streams = !towers_dual && e->side != nullptr && !e->single_stream;
  cudaStream_t s1 = two_streams ? e->side : s;
  if (two_streams) {
    VLA_CHECK_CUDA(cudaEventRecord(e->ev_fork, s));
    VLA_CHECK_CUDA(cudaStreamWaitEvent(e->side, e->ev_fork, 0));
  }
  int num_sms = 0, cur_dev = 0;
  cudaGetDevice(&cur_dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, cur_dev);
  const int tower_sms = (two_streams && e->split_sms && num_sms >= 8) ? (num_sms / 2) & ~1 : 0;
  int col_off = 0;
  if (towers_dual) {
    CK(vit_forward_both(e, s));
    for (int t = 0; t < 2; ++t) {
      const VitDims& v = e->vit[t];
      CK(copy_rows(e->va[t].x[v.used], v.dim, v.ntok, v.npre, e->feats + col_off, vd, P, 0, B, P, v.dim, s));
      col_off += v.dim;
    }
  } else {
    g_vla_sm_limit = tower_sms;
    for (int t = 0; t < 2; ++t) {
      cudaStream_t st = t == 0 ? s : s1;
      if (int rc = vit_forward(e, t, e->tr[two_streams ? t : 0], st)) {
        g_vla_sm_limit = 0;
        return rc;
      }
      const VitDims& v = e->vit[t];
      CK(copy_rows(e->va[t].x[v.used], v.dim, v.ntok, v.npre, e->feats + col_off, vd, P, 0, B, P, v.dim, st));
      col_off += v.dim;
    }
    g_vla_sm_limit = 0;
  }
  if (two_streams) {
    VLA_CHECK_CUDA(cudaEventRecord(e->ev_join, e->side));
    VLA_CHECK_CUDA(cudaStreamWaitEvent(s, e->ev_join, 0));
  }
  mark(1);
  {
    GemmEpilogue ep;
    ep.bias = e->pj_b[0];
    ep.act = 1;
    ep.preact_out = e->p1_pre;
    CK(G(e->feats, vd, e->pj_w[0], vd, e->tr[0].wide, phd, MP, phd, vd, ep, s));
    GemmEpilogue ep2;
    ep2.bias = e->pj_b[1];
    ep2.act = 1;
    ep2.preact_out = e->p2_pre;
    CK(G(e->tr[0].wide, phd, e->pj_w[1], phd, e->tr[0].a, h, MP, h, phd, ep2, s));
    GemmEpilogue ep3;   // fc3 writes straight into rows 1..P of each sample's multimodal sequence
    ep3.bias = e->pj_b[2];
    ep3.out_group = P;
    ep3.out_stride = L;
    ep3.out_offset = 1;
    CK(G(e->tr[0].a, h, e->pj_w[2], h, e->la.x[0], h, MP, h, h, ep3, s));
  }
  CK(embed_tokens_splice(e->ids, T, e->embed, e->la.x[0], B, T - 1, P, h, s));
  LlmActs& la = e->la;
  const bool prune = e->prune_last && e->fuse_swiglu_bwd && e->R > 0;
  for (int l = 0; l < c.llm_layers; ++l) {
    const LlamaLayerW& w = e->lw[l];
    CK(rmsnorm_fwd(la.x[l], w.n1, e->tr[0].norm, la.rstd1[l], ML, h, c.rms_eps, s));
    if (hd == 128) {   // RoPE fused into the q|k|v GEMM epilogue (pairs (c, c+64) of a head sit in one epilogue thread)
      GemmEpilogue ep;
      ep.pair_mode = 1;
      ep.rope_cos = e->rope_cos;
      ep.rope_sin = e->rope_sin;
      ep.rope_L = L;
      ep.rope_cols = 2 * h;
      CK(G(e->tr[0].norm, h, w.qkv, h, la.qkv[l], 3 * h, ML, 3 * h, h, ep, s));
    } else {
      CK(G(e->tr[0].norm, h, w.qkv, h, la.qkv[l], 3 * h, ML, 3 * h, h, plain, s));
      CK(rope_inplace(la.qkv[l], e->rope_cos, e->rope_sin, ML, L, NH, hd, +1, s));
    }
    CK(attention_fwd(la.qkv[l], la.attn_o[l], la.lse[l], e->kv_len, B, L, NH, hd, 1, s));
    if (prune && l == c.llm_layers - 1) {
      // Last layer: from here on every op is row-wise and only the R supervised rows of its output are read (lm_head on
      // the supervised rows), so gather those rows of the attention output / residual and run o_proj + MLP on R rows.
      const int R = e->R;
      CK(gather_rows(la.attn_o[l], e->sup_rows, e->ll_as, R, h, s));
      CK(gather_rows(la.x[l], e->sup_rows, e->ll_xs, R, h, s));
      GemmEpilogue ep;
      ep.resid = e->ll_xs;
      ep.ldr = h;
      CK(G(e->ll_as, h, w.o, h, e->ll_xm, h, R, h, h, ep, s));
      CK(rmsnorm_fwd(e->ll_xm, w.n2, e->ll_norm, e->ll_rstd2, R, h, c.rms_eps, s));
      GemmEpilogue eg;
      eg.pair_mode = 2;
      eg.act_out = e->ll_act;
      eg.ld_act = f;
      CK(G(e->ll_norm, h, w.gu, h, e->ll_gu, 2 * f, R, 2 * f, h, eg, s));
      GemmEpilogue ed;
      ed.resid = e->ll_xm;
      ed.ldr = h;
      CK(G(e->ll_act, f, w.down, f, e->hs, h, R, h, f, ed, s));   // = the supervised rows of the final hidden state
      break;
    }
    {
      GemmEpilogue ep;
      ep.resid = la.x[l];
      ep.ldr = h;
      CK(G(la.attn_o[l], h, w.o, h, la.x_mid[l], h, ML, h, h, ep, s));
    }
    CK(rmsnorm_fwd(la.x_mid[l], w.n2, e->tr[0].norm, la.rstd2[l], ML, h, c.rms_eps, s));
    {   // gate|up GEMM with the SwiGLU fused: raw gate|up saved for the backward, act = silu(gate)*up feeds down_proj
      GemmEpilogue ep;
      ep.pair_mode = 2;
      ep.act_out = e->tr[0].wide;
      ep.ld_act = f;
      CK(G(e->tr[0].norm, h, w.gu, h, la.gu[l], 2 * f, ML, 2 * f, h, ep, s));
    }
    {
      GemmEpilogue ep;
      ep.resid = la.x_mid[l];
      ep.ldr = h;
      CK(G(e->tr[0].wide, f, w.down, f, la.x[l + 1], h, ML, h, f, ep, s));
    }
  }
  const int R = e->R;
  if (!prune) CK(gather_rows(la.x[c.llm_layers], e->sup_rows, e->hs, R, h, s));
  CK(rmsnorm_fwd(e->hs, e->final_norm, e->hn, e->rstd_f, R, h, c.rms_eps, s));
  {
    GemmEpilogue ep;
    ep.out_f32 = 1;
    CK(G(e->hn, h, e->lm_head, h, e->logits, V, R, V, h, ep, s));
  }
  CK(loss_head_fwd_bwd(e->logits, e->meta, R, V, B, lp, e->row_stats, e->dlogits, scalars, pred_ids, s));
  // `action_preds = logits.argmax(dim=2)` of the reference's metrics (UADA.py:168,229; TMA.py:150,274): full vocabulary
  CK(argmax_rows(e->logits, R, V, e->pred_full, nullptr, 0, 0, s));
  e->last_pass_forward_only = (flags & VLA_FLAG_FORWARD_ONLY) != 0;
  if (flags & VLA_FLAG_FORWARD_ONLY) return 0;

  mark(2);
  // ---------------- backward (input gradients only) ----------------
  CK(G(e->dlogits, V, e->lm_head_t, V, e->tr[0].norm, h, R, h, V, plain, s));
  CK(rmsnorm_bwd(e->tr[0].norm, e->hs, e->final_norm, e->rstd_f, nullptr, e->hn, R, h, s));
  bf16* dx = e->tr[0].a;
  bf16* dxm = e->tr[0].b;
  int l_first = c.llm_layers - 1;
  if (prune) {
    // last layer on the R supervised rows: MLP backward, RMSNorm backward, o_proj backward; the attention backward then
    // spreads the gradient to every row through K and V, and the residual path carries d(x_mid) of the R rows only
    const int l = c.llm_layers - 1;
    const LlamaLayerW& w = e->lw[l];
    GemmEpilogue ep;
    ep.aux_mode = 2;
    ep.aux = e->ll_gu;
    ep.ldaux = 2 * f;
    CK(G(e->hn, h, w.down_t, h, e->ll_dgu, 2 * f, R, f, h, ep, s));
    CK(G(e->ll_dgu, 2 * f, w.gu_t, 2 * f, e->ll_dnorm, h, R, h, 2 * f, plain, s));
    CK(rmsnorm_bwd(e->ll_dnorm, e->ll_xm, w.n2, e->ll_rstd2, e->hn, e->ll_dxm, R, h, s));
    CK(G(e->ll_dxm, h, w.o_t, h, e->ll_dattn, h, R, h, h, plain, s));
    VLA_CHECK_CUDA(cudaMemsetAsync(e->tr[0].d, 0, static_cast<size_t>(ML) * h * sizeof(bf16), s));
    CK(scatter_rows(e->ll_dattn, e->sup_rows, e->tr[0].d, R, h, s));
    VLA_CHECK_CUDA(cudaMemsetAsync(dxm, 0, static_cast<size_t>(ML) * h * sizeof(bf16), s));
    CK(scatter_rows(e->ll_dxm, e->sup_rows, dxm, R, h, s));
    if (hd == 128) {
      CK(attention_bwd(la.qkv[l], la.attn_o[l], e->tr[0].d, la.lse[l], e->tr[0].delta, e->tr[0].qkv, e->kv_len, B, L, NH, hd, 1, e->rope_cos,
                       e->rope_sin, L, s));
    } else {
      CK(attention_bwd(la.qkv[l], la.attn_o[l], e->tr[0].d, la.lse[l], e->tr[0].delta, e->tr[0].qkv, e->kv_len, B, L, NH, hd, 1, nullptr, nullptr,
                       0, s));
      CK(rope_inplace(e->tr[0].qkv, e->rope_cos, e->rope_sin, ML, L, NH, hd, -1, s));
    }
    CK(G(e->tr[0].qkv, 3 * h, w.qkv_t, 3 * h, e->tr[0].norm, h, ML, h, 3 * h, plain, s));
    CK(rmsnorm_bwd(e->tr[0].norm, la.x[l], w.n1, la.rstd1[l], dxm, dx, ML, h, s));
    l_first = l - 1;
  } else {
    VLA_CHECK_CUDA(cudaMemsetAsync(dx, 0, static_cast<size_t>(ML) * h * sizeof(bf16), s));
    CK(scatter_rows(e->hn, e->sup_rows, dx, R, h, s));
  }
  const char* fd_env = getenv("VLA_FUSE_DELTA");   // read per step (A/B switch, tests)
  const bool fuse_delta = fd_env == nullptr || atoi(fd_env) != 0;
  for (int l = l_first; l >= 0; --l) {
    const LlamaLayerW& w = e->lw[l];
    if (e->fuse_swiglu_bwd) {   // d(act) = dX . W_down, with the SwiGLU backward fused: writes d(gate|up) directly
      GemmEpilogue ep;
      ep.aux_mode = 2;
      ep.aux = la.gu[l];
      ep.ldaux = 2 * f;
      CK(G(dx, h, w.down_t, h, e->tr[0].wide2, 2 * f, ML, f, h, ep, s));
    } else {
      CK(G(dx, h, w.down_t, h, e->tr[0].wide, f, ML, f, h, plain, s));
      CK(swiglu_bwd(e->tr[0].wide, la.gu[l], e->tr[0].wide2, ML, f, s));
    }
    CK(G(e->tr[0].wide2, 2 * f, w.gu_t, 2 * f, e->tr[0].norm, h, ML, h, 2 * f, plain, s));
    CK(rmsnorm_bwd(e->tr[0].norm, la.x_mid[l], w.n2, la.rstd2[l], dx, dxm, ML, h, s));
    // hd == 128: delta = rowsum(dO * O) comes out of the o_proj backward GEMM's epilogue (no pass of its own over dO and O)
    const bool fused_delta = fuse_delta && hd == 128 && attention_bwd_takes_delta(L, hd);
    if (fused_delta) {
      GemmEpilogue ep;
      ep.aux = la.attn_o[l];
      ep.ldaux = h;
      ep.delta_out = e->tr[0].delta;
      ep.delta_L = L;
      CK(G(dxm, h, w.o_t, h, e->tr[0].d, h, ML, h, h, ep, s));
    } else {
      CK(G(dxm, h, w.o_t, h, e->tr[0].d, h, ML, h, h, plain, s));
    }
    if (hd == 128) {   // RoPE backward fused into the attention backward's epilogues
      CK(attention_bwd(la.qkv[l], fused_delta ? nullptr : la.attn_o[l], e->tr[0].d, la.lse[l], e->tr[0].delta, e->tr[0].qkv, e->kv_len, B, L, NH,
                       hd, 1, e->rope_cos, e->rope_sin, L, s));
    } else {
      CK(attention_bwd(la.qkv[l], la.attn_o[l], e->tr[0].d, la.lse[l], e->tr[0].delta, e->tr[0].qkv, e->kv_len, B, L, NH, hd, 1, nullptr, nullptr,
                       0, s));
      CK(rope_inplace(e->tr[0].qkv, e->rope_cos, e->rope_sin, ML, L, NH, hd, -1, s));
    }
    CK(G(e->tr[0].qkv, 3 * h, w.qkv_t, 3 * h, e->tr[0].norm, h, ML, h, 3 * h, plain, s));
    CK(rmsnorm_bwd(e->tr[0].norm, la.x[l], w.n1, la.rstd1[l], dxm, dx, ML, h, s));
  }
  mark(3);
  // d X0 rows 1..P -> projector backward
  CK(copy_rows(dx, h, L, 1, e->tr[0].c, h, P, 0, B, P, h, s));
  {
    GemmEpilogue ep;   // GELU backward fused into the GEMM that produces d(post-activation)
    ep.aux_mode = 1;
    ep.aux = e->p2_pre;
    ep.ldaux = h;
    CK(G(e->tr[0].c, h, e->pj_t[2], h, e->tr[0].d, h, MP, h, h, ep, s));
    GemmEpilogue ep2;
    ep2.aux_mode = 1;
    ep2.aux = e->p1_pre;
    ep2.ldaux = phd;
    CK(G(e->tr[0].d, h, e->pj_t[1], h, e->tr[0].wide, phd, MP, phd, h, ep2, s));
  }
  CK(G(e->tr[0].wide, phd, e->pj_t[0], phd, e->feats, vd, MP, vd, phd, plain, s));
  if (two_streams) {
    VLA_CHECK_CUDA(cudaEventRecord(e->ev_fork, s));
    VLA_CHECK_CUDA(cudaStreamWaitEvent(e->side, e->ev_fork, 0));
  }
  col_off = 0;
  for (int t = 0; t < 2; ++t) {
    const VitDims& v = e->vit[t];
    const int64_t Mv = static_cast<int64_t>(B) * v.ntok;
    cudaStream_t st = t == 0 ? s : s1;
    Transients& tr = e->tr[(two_streams || towers_dual) ? t : 0];
    VLA_CHECK_CUDA(cudaMemsetAsync(tr.a, 0, static_cast<size_t>(Mv) * v.dim * sizeof(bf16), st));
    CK(copy_rows(e->feats + col_off, vd, P, 0, tr.a, v.dim, v.ntok, v.npre, B, P, v.dim, st));
    if (!towers_dual) {
      g_vla_sm_limit = tower_sms;
      const int rc_b = vit_backward(e, t, tr, st);
      g_vla_sm_limit = 0;
      if (rc_b) return rc_b;
    }
    col_off += v.dim;
  }
  if (towers_dual) CK(vit_backward_both(e, s));
  if (two_streams) {
    VLA_CHECK_CUDA(cudaEventRecord(e->ev_join, e->side));
    VLA_CHECK_CUDA(cudaStreamWaitEvent(s, e->ev_join, 0));
  }
  mark(4);
  CK(col2im_patches(e->va[0].a_col, e->va[1].a_col, e->dpx, B, H, H, c.patch, e->kpad, s));
  CK(patch_frontend_bwd(e->dpx, patch, xy, th, dpatch, B, H, H, ph, pw, fe_mode, e->nrm, s));
  mark(5);
  if (phase_timing) {
    cudaEventSynchronize(pev[5]);
    float t[5];
    for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&t[i], pev[i], pev[i + 1]);
    fprintf(stderr, "[phase ms] front+vit_fwd %.3f | projector+llm_fwd+loss %.3f | llm_bwd %.3f | projector+vit_bwd %.3f | tail %.3f\n", t[0],
            t[1], t[2], t[3], t[4]);
  }
  return 0;
}
}  // namespace

// =================================================================================================================
// vla_attack_step: the whole inner-loop body as one call / one CUDA graph (see include/vla_b200.h)
// =================================================================================================================
static long long g_graph_replays = 0;
extern "C" long long vla_graph_replays(void) { return g_graph_replays; }
extern "C" int vla_graph_kernel_nodes(const vla_engine* e) { return e ? e->last_kernel_nodes : 0; }

// Destroys the recorded attack-step graphs (they are re-recorded on demand).  Must precede vla_comm_destroy of a communicator
// whose all-reduce was recorded: NCCL keeps a communicator alive -- and blocks its destruction -- while graphs hold its work.
extern "C" int vla_engine_drop_graphs(vla_engine* e) {
  VLA_REQUIRE(e != nullptr, "vla_engine_drop_graphs: null engine");
  VLA_CHECK_CUDA(cudaDeviceSynchronize());
  for (auto& g : e->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
  return 0;
}

extern "C" int vla_engine_set_step_state(vla_engine* e, int placement_index, int adam_step, void* stream) {
  VLA_REQUIRE(e && e->ws, "vla_engine_set_step_state: call vla_engine_set_buffers first");
  VLA_REQUIRE(placement_index >= 0 && adam_step >= 0, "vla_engine_set_step_state: negative counter");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int v[2] = {placement_index, adam_step};   // passed by value into the copy below (cudaMemcpyAsync from pageable memory stages it)
  VLA_CHECK_CUDA(cudaMemcpyAsync(e->dstate, v, sizeof(v), cudaMemcpyHostToDevice, s));
  e->h_place = placement_index;
  e->h_adam = adam_step;
  return 0;
}

extern "C" int vla_engine_get_step_state(const vla_engine* e, int* placement_index, int* adam_step) {
  VLA_REQUIRE(e != nullptr, "vla_engine_get_step_state: null engine");
  if (placement_index) *placement_index = e->h_place;
  if (adam_step) *adam_step = e->h_adam;
  return 0;
}

namespace {

// The launch sequence of one attack iteration on stream s (eager or being captured).
int attack_step_sequence(vla_engine* e, float* patch, float* m, float* v, float* dpatch, float* acc, const vla_step_params* sp,
                         vla_comm* comm, float* scal_hist, int* pred_ids, cudaStream_t s) {
  const int n = 3 * sp->ph * sp->pw;
  const bool update = !(sp->flags & VLA_STEP_NO_UPDATE);
  CK(step_begin(e->dstate, e->xy, e->theta, e->xy_cur, e->theta_cur, e->B, e->n_place > 0 ? e->n_place : 1, e->scal_cur, s));
  CK(fwd_bwd_impl(e, patch, sp->ph, sp->pw, e->xy_cur, e->theta_cur, sp->fe_mode, &sp->loss, dpatch, e->scal_cur, pred_ids, 0, s));
  float* g = dpatch;
  if (acc) {   // TMA / UPA with accumulate_steps > 1: gradients pile up over outer iterations (TMA.py:162-170)
    CK(accumulate_f32(acc, dpatch, n, s));
    g = acc;
  }
  const int world = vla_comm_world(comm);
  if (comm && world > 1) CK(vla_allreduce_patch_grad(comm, g, n, s));   // DDP reducer (UADA_ddp.py:206)
  if (update)
    CK(patch_update_dev(patch, g, m, v, n, e->dstate, sp->beta1, sp->beta2, sp->eps, sp->opt_kind, 1.f / static_cast<float>(world),
                        sp->clip_l1, e->scal_cur, acc, s));
  CK(step_end(e->dstate, e->scal_cur, scal_hist, update ? 1 : 0, s));
  return 0;
}

struct StepKey {   // everything the recorded launch sequence depends on
  int B, T, R, n_place;
  vla_step_params sp;
  const void *patch, *m, *v, *dpatch, *acc, *comm, *hist, *pred;
  int single_stream, attn_impl, pad;
};

}  // namespace

extern "C" int vla_attack_step(vla_engine* e, float* patch, float* exp_avg, float* exp_avg_sq, float* dpatch, float* accumulate,
                               const vla_step_params* sp, vla_comm* comm, float* scalars_hist, int* pred_ids, void* stream) {
  VLA_REQUIRE(e && e->weights_resolved, "vla_attack_step: engine not initialised");
  VLA_REQUIRE(e->batch_set, "vla_attack_step: call vla_engine_set_batch first");
  VLA_REQUIRE(e->rope_set, "vla_attack_step: call vla_engine_set_rope first");
  VLA_REQUIRE(patch && dpatch && sp && pred_ids, "vla_attack_step: null argument");
  VLA_REQUIRE(sp->opt_kind == OPT_ADAMW || sp->opt_kind == OPT_PGD, "vla_attack_step: bad optimiser kind %d", sp->opt_kind);
  VLA_REQUIRE(sp->opt_kind != OPT_ADAMW || (exp_avg && exp_avg_sq) || (sp->flags & VLA_STEP_NO_UPDATE),
              "vla_attack_step: AdamW needs the moment buffers");
  VLA_REQUIRE(sp->fe_mode == FE_MODE_NONE || e->h_place < e->n_place,
              "vla_attack_step: placement %d not uploaded (have %d): call vla_engine_set_placements / vla_engine_set_step_state",
              e->h_place, e->n_place);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // learning rate -> device state (changes once per outer iteration: LambdaLR is stepped outside the inner loop, UADA.py:162-164)
  if (sp->lr != e->h_lr) {
    if (!e->lr_ring) VLA_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&e->lr_ring), 64 * sizeof(float), cudaHostAllocDefault));
    float* slot = e->lr_ring + (e->lr_ring_pos++ & 63);
    *slot = sp->lr;
    VLA_CHECK_CUDA(cudaMemcpyAsync(&e->dstate->lr, slot, sizeof(float), cudaMemcpyHostToDevice, s));
    e->h_lr = sp->lr;
  }
  const bool update = !(sp->flags & VLA_STEP_NO_UPDATE);
  static const bool graphs_off = getenv("VLA_STEP_GRAPH") && atoi(getenv("VLA_STEP_GRAPH")) == 0;
  const bool eager = (sp->flags & VLA_STEP_NO_GRAPH) || graphs_off || phase_timing_on() || vla_gemm_profiling();
  int rc = 0;
  if (eager) {
    rc = attack_step_sequence(e, patch, exp_avg, exp_avg_sq, dpatch, accumulate, sp, comm, scalars_hist, pred_ids, s);
  } else {
    StepKey k;
    memset(&k, 0, sizeof(k));
    k.B = e->B; k.T = e->T; k.R = e->R; k.n_place = e->n_place;
    k.sp = *sp;
    k.sp.lr = 0.f;   // read from device memory
    k.patch = patch; k.m = exp_avg; k.v = exp_avg_sq; k.dpatch = dpatch; k.acc = accumulate; k.comm = comm; k.hist = scalars_hist;
    k.pred = pred_ids;
    k.single_stream = e->single_stream ? 1 : 0;
    k.attn_impl = attention_impl();
    vla_engine::StepGraph* g = nullptr;
    for (auto& cand : e->graphs)
      if (cand.key.size() == sizeof(k) && memcmp(cand.key.data(), &k, sizeof(k)) == 0) g = &cand;
    if (!g) {
      if (e->graphs.size() >= 16) {   // bounded cache: drop the oldest recording
        if (e->graphs.front().exec) cudaGraphExecDestroy(e->graphs.front().exec);
        e->graphs.erase(e->graphs.begin());
      }
      e->graphs.emplace_back();
      g = &e->graphs.back();
      g->key.assign(reinterpret_cast<uint8_t*>(&k), reinterpret_cast<uint8_t*>(&k) + sizeof(k));
    }
    g->calls++;
    if (g->calls == 1) {
      // first sight: eager (the GEMM autotuner times its variants on first use of a shape, which cannot happen inside a capture)
      rc = attack_step_sequence(e, patch, exp_avg, exp_avg_sq, dpatch, accumulate, sp, comm, scalars_hist, pred_ids, s);
    } else {
      if (!g->exec) {
        if (!e->cap) VLA_CHECK_CUDA(cudaStreamCreateWithFlags(&e->cap, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        VLA_CHECK_CUDA(cudaStreamBeginCapture(e->cap, cudaStreamCaptureModeThreadLocal));
        const long long launches0 = g_vla_launch_count;
        rc = attack_step_sequence(e, patch, exp_avg, exp_avg_sq, dpatch, accumulate, sp, comm, scalars_hist, pred_ids, e->cap);
        const cudaError_t ce = cudaStreamEndCapture(e->cap, &graph);
        g_vla_launch_count = launches0;   // nothing ran: the replays are counted below
        if (rc != 0 || ce != cudaSuccess || !graph) {
          if (graph) cudaGraphDestroy(graph);
          if (rc == 0) vla_set_error("vla_attack_step: stream capture failed: %s", cudaGetErrorString(ce));
          cudaGetLastError();
          g->calls = 0;
          return rc ? rc : 1;
        }
        size_t nn = 0;
        VLA_CHECK_CUDA(cudaGraphGetNodes(graph, nullptr, &nn));
        std::vector<cudaGraphNode_t> nodes(nn);
        VLA_CHECK_CUDA(cudaGraphGetNodes(graph, nodes.data(), &nn));
        int kn = 0;
        for (size_t i = 0; i < nn; ++i) {
          cudaGraphNodeType ty;
          if (cudaGraphNodeGetType(nodes[i], &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel) ++kn;
        }
        const cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
          vla_set_error("vla_attack_step: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
          g->exec = nullptr;
          g->calls = 0;
          return 1;
        }
        g->kernel_nodes = kn;
      }
      VLA_CHECK_CUDA(cudaGraphLaunch(g->exec, s));
      e->last_kernel_nodes = g->kernel_nodes;
      g_vla_launch_count += g->kernel_nodes;   // the recorded kernels run on every replay
      ++g_graph_replays;
    }
  }
  if (rc) return rc;
  e->h_place++;
  if (update) e->h_adam++;
  return 0;
}

// =================================================================================================================
// Greedy action decode with the KV cache of the last forward-only pass (predict_action, modeling_prismatic.py:506-536 ->
// HF generate(do_sample=False, max_new_tokens=action_dim)).  See decode.cu.
// =================================================================================================================
extern "C" int vla_engine_decode_greedy(vla_engine* e, int prompt_len, int n_tokens, int* tokens, void* stream) {
  VLA_REQUIRE(e && e->weights_resolved && e->batch_set, "vla_engine_decode_greedy: engine / batch not set");
  VLA_REQUIRE(tokens && n_tokens >= 1, "vla_engine_decode_greedy: null / empty output");
  VLA_REQUIRE(e->last_pass_forward_only, "vla_engine_decode_greedy: run the prefill first (vla_fwd_bwd with VLA_FLAG_FORWARD_ONLY)");
  const vla_config& c = e->cfg;
  const int B = e->B, L = e->L, P = e->np, h = c.llm_hidden, f = c.llm_ffn, V = c.vocab, NH = c.llm_heads, hd = h / NH;
  VLA_REQUIRE(e->R == B, "vla_engine_decode_greedy: the prefill must have exactly one supervised row per sample (got %d rows for %d samples)", e->R, B);
  for (int b = 0; b < B; ++b)
    VLA_REQUIRE(e->h_sup_rows[b] == b * L + P + prompt_len - 1,
                "vla_engine_decode_greedy: sample %d: the supervised row must be the last prompt position (prompts of equal length %d)", b, prompt_len);
  VLA_REQUIRE(P + prompt_len + n_tokens - 2 <= L - 1, "vla_engine_decode_greedy: plan too short: T must be >= prompt_len + n_tokens (= %d)",
              prompt_len + n_tokens);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  constexpr int DEC_MAX_TOKENS = 32;
  VLA_REQUIRE(n_tokens <= DEC_MAX_TOKENS, "vla_engine_decode_greedy: at most %d tokens", DEC_MAX_TOKENS);
  // M = B projections: the HBM-bound skinny kernels for B <= 4 (the evaluation loop runs B = 1), the tcgen05 GEMM otherwise
  const bool no_gemv = getenv("VLA_DECODE_GEMV") && atoi(getenv("VLA_DECODE_GEMV")) == 0;     // read per call: A/B switches
  const bool no_graph = getenv("VLA_DECODE_GRAPH") && atoi(getenv("VLA_DECODE_GRAPH")) == 0;
  const bool skinny = !no_gemv && gemv_supported(B, h, h, h) && gemv_supported(B, f, f, f) && f % 64 == 0;
  bf16* x = e->ll_xs;
  // token 0: the argmax of the prefill's logits row (full vocabulary, as generate() does), and its embedding
  if (skinny) {
    CK(decode_select(e->logits, B, V, e->dec_ids, e->dec_tokens, DEC_MAX_TOKENS, 0, nullptr, e->embed, x, h, s));
  } else {
    CK(argmax_rows(e->logits, B, V, e->dec_ids, e->dec_tokens, DEC_MAX_TOKENS, 0, s));
    CK(embed_rows(e->dec_ids, e->embed, x, B, h, s));
  }
  // One decode step: x holds the embeddings of the previous token on entry and of the new token on exit.
  // Skinny path, five kernels per layer (decode.cu); ds != nullptr: every position-dependent kernel reads the position from
  // device memory, so the step can be recorded once per plan and replayed for every token of every action.
  auto decode_step_skinny = [&](int pos, int k, int* ds, cudaStream_t st) -> int {
    for (int l = 0; l < c.llm_layers; ++l) {
      const LlamaLayerW& w = e->lw[l];
      // RMSNorm + q|k|v of the new position, written straight into the layer's cache row b * L + pos
      CK(gemv_bf16(x, h, w.n1, c.rms_eps, w.qkv, h, e->la.qkv[l], 3 * h, B, 3 * h, h, nullptr, 0, 0, 0, L, pos, ds, st));
      CK(attention_decode(e->la.qkv[l], e->ll_as, e->rope_cos, e->rope_sin, B, L, pos, NH, hd, ds, st));
      CK(gemv_bf16(e->ll_as, h, nullptr, 0.f, w.o, h, e->ll_xm, h, B, h, h, x, h, 0, 0, 0, 0, nullptr, st));
      CK(gemv_bf16(e->ll_xm, h, w.n2, c.rms_eps, w.gu, h, e->ll_act, f, B, 2 * f, h, nullptr, 0, 0, 1, 0, 0, nullptr, st));
      CK(gemv_bf16(e->ll_act, f, nullptr, 0.f, w.down, f, x, h, B, h, f, e->ll_xm, h, 0, 0, 0, 0, nullptr, st));
    }
    CK(gemv_bf16(x, h, e->final_norm, c.rms_eps, e->lm_head, h, e->logits, V, B, V, h, nullptr, 0, 1, 0, 0, 0, nullptr, st));
    CK(decode_select(e->logits, B, V, e->dec_ids, e->dec_tokens, DEC_MAX_TOKENS, k, ds, e->embed, x, h, st));
    return 0;
  };
  // B > 4: the same step on the tcgen05 GEMM (M = B), norms and SwiGLU as in the prefill
  auto decode_step_gemm = [&](int pos, int k, cudaStream_t st) -> int {
    auto linear = [&](const bf16* A, int K, const bf16* W, void* out, int64_t ldc, int N, const bf16* resid, int out_f32, int remap_stride) -> int {
      GemmEpilogue ep;
      ep.resid = resid;
      ep.ldr = N;
      ep.out_f32 = out_f32;
      if (remap_stride) {
        ep.out_group = 1;
        ep.out_stride = remap_stride;
        ep.out_offset = pos;
      }
      return G(A, K, W, K, out, ldc, B, N, K, ep, st);
    };
    for (int l = 0; l < c.llm_layers; ++l) {
      const LlamaLayerW& w = e->lw[l];
      CK(rmsnorm_fwd(x, w.n1, e->ll_norm, e->ll_rstd2, B, h, c.rms_eps, st));
      CK(linear(e->ll_norm, h, w.qkv, e->la.qkv[l], 3 * h, 3 * h, nullptr, 0, L));
      CK(attention_decode(e->la.qkv[l], e->ll_as, e->rope_cos, e->rope_sin, B, L, pos, NH, hd, nullptr, st));
      CK(linear(e->ll_as, h, w.o, e->ll_xm, h, h, x, 0, 0));
      CK(rmsnorm_fwd(e->ll_xm, w.n2, e->ll_norm, e->ll_rstd2, B, h, c.rms_eps, st));
      GemmEpilogue ep;
      ep.pair_mode = 2;
      ep.act_out = e->ll_act;
      ep.ld_act = f;
      CK(G(e->ll_norm, h, w.gu, h, e->ll_gu, 2 * f, B, 2 * f, h, ep, st));
      CK(linear(e->ll_act, f, w.down, x, h, h, e->ll_xm, 0, 0));
    }
    CK(rmsnorm_fwd(x, e->final_norm, e->hn, e->rstd_f, B, h, c.rms_eps, st));
    CK(linear(e->hn, h, e->lm_head, e->logits, V, V, nullptr, 1, 0));
    CK(argmax_rows(e->logits, B, V, e->dec_ids, e->dec_tokens, DEC_MAX_TOKENS, k, st));
    CK(embed_rows(e->dec_ids, e->embed, x, B, h, st));
    return 0;
  };
  const int pos0 = P + prompt_len;   // cache row (within a sample) of the first generated token
  if (skinny && !no_graph && n_tokens > 1) {
    const int init[2] = {pos0, 1};
    VLA_CHECK_CUDA(cudaMemcpyAsync(e->dec_state, init, sizeof(init), cudaMemcpyHostToDevice, s));
    if (!e->dec_graph) {   // record one step (nothing position-dependent is baked in; no autotuned kernels on this path)
      if (!e->cap) VLA_CHECK_CUDA(cudaStreamCreateWithFlags(&e->cap, cudaStreamNonBlocking));
      cudaGraph_t graph = nullptr;
      VLA_CHECK_CUDA(cudaStreamBeginCapture(e->cap, cudaStreamCaptureModeThreadLocal));
      const long long launches0 = g_vla_launch_count;
      const int rc = decode_step_skinny(0, 0, e->dec_state, e->cap);
      const cudaError_t ce = cudaStreamEndCapture(e->cap, &graph);
      g_vla_launch_count = launches0;
      if (rc != 0 || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        if (rc == 0) vla_set_error("vla_engine_decode_greedy: stream capture failed: %s", cudaGetErrorString(ce));
        cudaGetLastError();
        return rc ? rc : 1;
      }
      const cudaError_t ie = cudaGraphInstantiate(&e->dec_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) {
        e->dec_graph = nullptr;
        vla_set_error("vla_engine_decode_greedy: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
        return 1;
      }
    }
    for (int k = 1; k < n_tokens; ++k) VLA_CHECK_CUDA(cudaGraphLaunch(e->dec_graph, s));
  } else {
    for (int k = 1; k < n_tokens; ++k) {
      if (skinny) CK(decode_step_skinny(pos0 + k - 1, k, nullptr, s));
      else CK(decode_step_gemm(pos0 + k - 1, k, s));
    }
  }
  VLA_CHECK_CUDA(cudaMemcpy2DAsync(tokens, sizeof(int) * n_tokens, e->dec_tokens, sizeof(int) * DEC_MAX_TOKENS, sizeof(int) * n_tokens, B,
                                   cudaMemcpyDeviceToDevice, s));
  return 0;
}

// Debug / test taps: copy an internal activation to a caller buffer (device). Returns the element count.
extern "C" int64_t vla_engine_debug_tap(vla_engine* e, const char* what_c, void* dst, int64_t max_bytes, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::string what(what_c);
  const void* src = nullptr;
  int64_t bytes = 0, elems = 0;
  const vla_config& c = e->cfg;
  auto set = [&](const void* p, int64_t n, int esz) { src = p; elems = n; bytes = n * esz; };
  const int64_t ML = static_cast<int64_t>(e->B) * e->L;
  if (what == "px") set(e->px, static_cast<int64_t>(e->B) * 6 * c.img * c.img, 2);
  else if (what == "dpx") set(e->dpx, static_cast<int64_t>(e->B) * 6 * c.img * c.img, 2);
  else if (what == "dino_out") set(e->va[0].x[e->vit[0].used], static_cast<int64_t>(e->B) * e->vit[0].ntok * e->vit[0].dim, 2);
  else if (what == "siglip_out") set(e->va[1].x[e->vit[1].used], static_cast<int64_t>(e->B) * e->vit[1].ntok * e->vit[1].dim, 2);
  else if (what == "dino_x0") set(e->va[0].x[0], static_cast<int64_t>(e->B) * e->vit[0].ntok * e->vit[0].dim, 2);
  else if (what == "dino_x1") set(e->va[0].x[1], static_cast<int64_t>(e->B) * e->vit[0].ntok * e->vit[0].dim, 2);
  else if (what == "llm_x0") set(e->la.x[0], ML * c.llm_hidden, 2);
  else if (what == "llm_x1") set(e->la.x[1], ML * c.llm_hidden, 2);
  else if (what == "llm_out") set(e->la.x[c.llm_layers], ML * c.llm_hidden, 2);
  else if (what == "logits") set(e->logits, static_cast<int64_t>(e->R) * c.vocab, 4);
  else if (what == "dlogits") set(e->dlogits, static_cast<int64_t>(e->R) * c.vocab, 2);
  else if (what == "d_llm_x0") set(e->tr[0].a, ML * c.llm_hidden, 2);
  else {
    vla_set_error("unknown tap '%s'", what_c);
    return -1;
  }
  if (bytes > max_bytes) {
    vla_set_error("tap '%s' needs %lld bytes, buffer has %lld", what_c, (long long)bytes, (long long)max_bytes);
    return -1;
  }
  if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
    vla_set_error("tap copy failed");
    return -1;
  }
  return elems;
}
