// Thin torch C++ extension over the C ABI (include/ens_render.h): the autograd plumbing of the hot calls in C++.
//
// The reference's callers (src/Tracker.py:159-201, src/Mapper.py:523-578) run EAGERLY: every iteration calls get_samples per
// keyframe, Renderer.render_batch_ray once and loss.backward().  With the kernels at ~0.5 ms per iteration, a Python
// torch.autograd.Function that takes ~75 tensors (3 grids + 69 decoder parameters), marshals them through ctypes and hands
// back ~75 gradient views costs several times the GPU work.  This file does exactly that plumbing -- allocate outputs /
// saved-for-backward / gradient arena, fill EnsScene / EnsGrads, call ens_render_fwd / ens_render_bwd, slice the flat decoder
// gradients into per-parameter views -- inside torch::autograd::Function, so the Python drop-ins (functional.py) only
// forward their arguments.  No arithmetic lives here.
//
// Built in-tree by __graft_entry__.build() (torch.utils.cpp_extension) and linked against libens_render.so.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "../../include/ens_render.h"

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

namespace {

int64_t g_launches = 0;          // C-ABI calls made through this extension (bench.py's gpu_launches counts them)

void check(int rc, const char *what) {
  if (rc == ENS_OK) return;
  std::string msg = ens_strerror(rc);
  if (rc == ENS_ECUDA) msg += std::string(": ") + ens_last_error();
  TORCH_CHECK(false, "libens_render: ", what, " failed: ", msg, " (code ", rc, ")");
}

inline void *ptr(const Tensor &t) { return t.defined() && t.numel() > 0 ? t.data_ptr() : nullptr; }
inline ens_stream_t cur_stream(const Tensor &t) { return (ens_stream_t)at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }
inline Tensor f32c(const Tensor &t) {
  Tensor x = t.detach();
  if (x.scalar_type() != torch::kFloat32) x = x.to(torch::kFloat32);
  return x.is_contiguous() ? x : x.contiguous();
}

const int kLevelsOfStage[4][3] = {{0, -1, -1}, {1, -1, -1}, {1, 2, -1}, {1, 2, 3}};

// a [1,32,Z,Y,X] view of [Z,Y,X,32] storage (scene.as_native_layout)?
bool is_native_strided(const Tensor &g) {
  if (g.dim() != 5) return false;
  const auto s = g.strides();
  const auto z = g.sizes();
  return s[1] == 1 && s[4] == 32 && s[3] == 32 * z[4] && s[2] == 32 * z[4] * z[3];
}

EnsScene make_scene(int stage, const std::vector<double> &bounds, const std::vector<Tensor> &native, const std::vector<Tensor> &packed) {
  EnsScene sc;
  std::memset(&sc, 0, sizeof(sc));
  int k = 0;
  for (int q = 0; q < 3; ++q) {
    const int lv = kLevelsOfStage[stage][q];
    if (lv < 0) continue;
    const Tensor &n = native[k];
    sc.grid[lv] = (const float *)n.data_ptr();
    sc.dims[lv][0] = (int32_t)n.size(0); sc.dims[lv][1] = (int32_t)n.size(1); sc.dims[lv][2] = (int32_t)n.size(2);
    sc.weights[lv] = (const float *)packed[k].data_ptr();
    ++k;
  }
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 2; ++b) { sc.bound[a][b] = bounds[a * 2 + b]; sc.coarse_bound[a][b] = bounds[6 + a * 2 + b]; }
  return sc;
}

// ---------------------------------------------------------------------------------------------
// Renderer.render_batch_ray (src/utils/Renderer.py:64-199)
//   tensor inputs (autograd edges, in this order): rays_o, rays_d, grids (n_levels), params (sum of the levels' tensors),
//   then non-differentiable helpers: aux = {gt_depth | empty, depth_max | empty, t_vals, t_surf | empty}, native, packed
// ---------------------------------------------------------------------------------------------
struct RenderFn : public torch::autograd::Function<RenderFn> {
  static variable_list forward(AutogradContext *ctx, const Tensor &rays_o, const Tensor &rays_d, at::TensorList grids,
                               at::TensorList params, at::TensorList aux, at::TensorList native, at::TensorList packed,
                               int64_t stage, int64_t n_samples, int64_t n_surface, std::vector<double> bounds,
                               std::vector<int64_t> n_params, bool want_aux, bool grad_enabled, bool tc_map, bool tc_pose) {
    const c10::cuda::CUDAGuard guard(rays_o.device());
    const Tensor ro = f32c(rays_o), rd = f32c(rays_d);
    const int64_t R = ro.size(0);
    const bool has_depth = aux[0].numel() > 0 && stage != ENS_STAGE_COARSE;
    const int S = (int)(n_samples + (has_depth ? n_surface : 0));
    const auto dev = ro.device();
    const auto f64 = torch::TensorOptions().dtype(torch::kFloat64).device(dev);
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
    Tensor depth = torch::empty({R}, f64), var = torch::empty({R}, f64), color = torch::empty({R, 3}, f32);
    Tensor raw = torch::empty({R, S, 4}, f32);
    Tensor z, w;
    if (want_aux) { z = torch::empty({R, S}, f64); w = torch::empty({R, S}, f32); }
    Tensor gd;
    if (has_depth) gd = f32c(aux[0]).reshape({-1});

    bool want_grid = false, want_dec = false;
    const bool want_rays = grad_enabled && (rays_o.requires_grad() || rays_d.requires_grad());
    if (grad_enabled) {
      for (const Tensor &g : grids) want_grid |= g.requires_grad();
      for (const Tensor &p : params) want_dec |= p.requires_grad();
    }
    const bool want_bwd = want_rays || want_grid || want_dec;
    std::vector<Tensor> nat(native.begin(), native.end()), pk(packed.begin(), packed.end());
    EnsScene sc = make_scene((int)stage, bounds, nat, pk);
    EnsRenderCfg cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.n_samples = (int32_t)n_samples; cfg.n_surface = (int32_t)n_surface; cfg.occupancy = 1;
    cfg.t_vals = (const float *)aux[2].data_ptr();
    cfg.t_vals_surface = aux[3].numel() > 0 ? (const double *)aux[3].data_ptr() : nullptr;

    // saved-for-backward state (what autograd keeps as saved tensors in the reference); kinds as in ens_render.h
    Tensor saved, scratch;
    int kind = want_dec ? 1 : 0;
    bool map_tc = want_dec && tc_map && stage != ENS_STAGE_COARSE;
    if (map_tc) {
      const int64_t nb = ens_fwd_saved_bytes_kind(R, S, (int)stage, 3);
      if (nb > 0) { saved = torch::empty({nb / 4}, torch::TensorOptions().dtype(torch::kInt32).device(dev)); kind = 3; }
      else map_tc = false;
    }
    if (want_bwd && !map_tc) {
      const int64_t nb = ens_fwd_saved_bytes(R, S, (int)stage, want_dec ? 1 : 0);
      if (nb > 0) saved = torch::empty({nb / 4}, torch::TensorOptions().dtype(torch::kInt32).device(dev));
    }
    const bool tc_forward = !want_bwd || map_tc || (saved.defined() && !want_dec && tc_pose);
    if (tc_forward) {
      const int64_t nb = ens_fwd_scratch_bytes(R, S, (int)stage);
      if (nb > 0) {
        scratch = torch::empty({nb / 8}, f64);
        if (saved.defined() && !map_tc) kind = 2;
      }
    }
    ++g_launches;
    check(ens_render_fwd(&sc, &cfg, (int)stage, (const float *)ro.data_ptr(), (const float *)rd.data_ptr(),
                         has_depth ? (const float *)gd.data_ptr() : nullptr, has_depth ? (const double *)aux[1].data_ptr() : nullptr, R,
                         (double *)depth.data_ptr(), (double *)var.data_ptr(), (float *)color.data_ptr(),
                         want_aux ? (double *)z.data_ptr() : nullptr, want_aux ? (float *)w.data_ptr() : nullptr, (float *)raw.data_ptr(),
                         ptr(saved), saved.defined() ? saved.numel() * 4 : 0, kind, ptr(scratch),
                         scratch.defined() ? scratch.numel() * 8 : 0, cur_stream(ro)),
          "ens_render_fwd");

    if (want_bwd) {
      variable_list keep = {ro, rd, has_depth ? gd : Tensor(), has_depth ? aux[1] : Tensor(), raw, saved.defined() ? saved : Tensor(),
                            aux[2], aux[3]};
      for (const Tensor &t : nat) keep.push_back(t);
      for (const Tensor &t : pk) keep.push_back(t);
      ctx->save_for_backward(keep);
      ctx->saved_data["stage"] = stage; ctx->saved_data["n_samples"] = n_samples; ctx->saved_data["n_surface"] = n_surface;
      ctx->saved_data["S"] = (int64_t)S; ctx->saved_data["has_depth"] = has_depth; ctx->saved_data["kind"] = (int64_t)kind;
      ctx->saved_data["bounds"] = bounds; ctx->saved_data["n_params"] = n_params;
      ctx->saved_data["n_levels"] = (int64_t)nat.size();
      std::vector<int64_t> gns;
      for (const Tensor &g : grids) gns.push_back(is_native_strided(g) ? 1 : 0);
      ctx->saved_data["grid_native"] = gns;
      // reference shapes of the parameters (views of the flat gradient)
      std::vector<std::vector<int64_t>> shapes;
      for (const Tensor &p : params) shapes.push_back(p.sizes().vec());
      ctx->saved_data["param_shapes"] = shapes;
      ctx->saved_data["want_dec"] = want_dec;
    }
    variable_list out = {depth, var, color, raw};
    if (want_aux) { out.push_back(z); out.push_back(w); }
    if (want_aux) ctx->mark_non_differentiable({raw, z, w});
    else ctx->mark_non_differentiable({raw});
    return out;
  }

  static variable_list backward(AutogradContext *ctx, variable_list gout) {
    const variable_list sv = ctx->get_saved_variables();
    const Tensor &ro = sv[0], &rd = sv[1], &gd = sv[2], &dmax = sv[3], &raw = sv[4], &saved = sv[5], &t_vals = sv[6], &t_surf = sv[7];
    const c10::cuda::CUDAGuard guard(ro.device());
    const int64_t stage = ctx->saved_data["stage"].toInt(), S = ctx->saved_data["S"].toInt();
    const int64_t n_levels = ctx->saved_data["n_levels"].toInt();
    const bool has_depth = ctx->saved_data["has_depth"].toBool();
    int kind = (int)ctx->saved_data["kind"].toInt();
    const std::vector<double> bounds = ctx->saved_data["bounds"].toDoubleVector();
    const std::vector<int64_t> n_params = ctx->saved_data["n_params"].toIntVector();
    const std::vector<int64_t> grid_native = ctx->saved_data["grid_native"].toIntVector();
    const bool had_dec = ctx->saved_data["want_dec"].toBool();
    std::vector<Tensor> nat(sv.begin() + 8, sv.begin() + 8 + n_levels), pk(sv.begin() + 8 + n_levels, sv.begin() + 8 + 2 * n_levels);
    const int64_t R = ro.size(0);
    const auto dev = ro.device();
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);

    // autograd edges: 0 rays_o, 1 rays_d, 2.. grids, then params
    const bool need_ro = ctx->needs_input_grad(0), need_rd = ctx->needs_input_grad(1);
    int64_t total_params = 0;
    for (int64_t n : n_params) total_params += n;
    bool any_param = false;
    for (int64_t i = 0; i < total_params; ++i) any_param |= ctx->needs_input_grad(2 + n_levels + i);
    any_param = any_param && had_dec;

    // ONE zero-filled arena for every accumulate-into sink (native grid gradients, flat decoder gradients)
    std::vector<int64_t> g_off(n_levels, -1), d_off(n_levels, -1), d_len(n_levels, 0);
    int64_t pos = 0;
    int lv_of[3];
    { int k = 0; for (int q = 0; q < 3; ++q) if (kLevelsOfStage[stage][q] >= 0) lv_of[k++] = kLevelsOfStage[stage][q]; }
    for (int64_t k = 0; k < n_levels; ++k)
      if (ctx->needs_input_grad(2 + k)) { g_off[k] = pos; pos += (nat[k].numel() + 3) & ~(int64_t)3; }
    if (any_param)
      for (int64_t k = 0; k < n_levels; ++k) { d_len[k] = ens_decoder_grad_floats(lv_of[k]); d_off[k] = pos; pos += (d_len[k] + 3) & ~(int64_t)3; }
    Tensor arena;
    if (pos > 0) arena = torch::zeros({pos}, f32);
    EnsGrads grads;
    std::memset(&grads, 0, sizeof(grads));
    for (int64_t k = 0; k < n_levels; ++k) {
      if (g_off[k] >= 0) grads.grid[lv_of[k]] = (float *)arena.data_ptr() + g_off[k];
      if (d_off[k] >= 0) grads.decoder[lv_of[k]] = (float *)arena.data_ptr() + d_off[k];
    }
    Tensor g_ro, g_rd;
    if (need_ro || need_rd) { g_ro = torch::empty_like(ro); g_rd = torch::empty_like(rd); }
    grads.rays_o = (float *)ptr(g_ro); grads.rays_d = (float *)ptr(g_rd);
    Tensor sv_used = saved;
    if (sv_used.defined() && any_param && !(kind == 1 || kind == 3)) sv_used = Tensor();      // activations were not kept: recompute
    const int64_t ws_bytes = ens_bwd_workspace_bytes(R, (int)S, any_param ? 1 : 0);
    Tensor ws;
    if (ws_bytes > 0) ws = torch::empty({(ws_bytes + 3) / 4}, f32);
    EnsScene sc = make_scene((int)stage, bounds, nat, pk);
    EnsRenderCfg cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.n_samples = (int32_t)ctx->saved_data["n_samples"].toInt(); cfg.n_surface = (int32_t)ctx->saved_data["n_surface"].toInt();
    cfg.occupancy = 1;
    cfg.t_vals = (const float *)t_vals.data_ptr();
    cfg.t_vals_surface = t_surf.defined() && t_surf.numel() > 0 ? (const double *)t_surf.data_ptr() : nullptr;
    Tensor gdp, gvp, gcp;
    if (gout[0].defined()) gdp = gout[0].detach().to(torch::kFloat64).contiguous();
    if (gout[1].defined()) gvp = gout[1].detach().to(torch::kFloat64).contiguous();
    if (gout[2].defined()) gcp = f32c(gout[2]);
    ++g_launches;
    check(ens_render_bwd(&sc, &cfg, (int)stage, (const float *)ro.data_ptr(), (const float *)rd.data_ptr(),
                         has_depth ? (const float *)gd.data_ptr() : nullptr, has_depth ? (const double *)dmax.data_ptr() : nullptr, R,
                         (const float *)raw.data_ptr(), (const double *)ptr(gdp), (const double *)ptr(gvp), (const float *)ptr(gcp), &grads,
                         ptr(ws), ws_bytes, ptr(sv_used), sv_used.defined() ? sv_used.numel() * 4 : 0, sv_used.defined() ? kind : 0,
                         cur_stream(ro)),
          "ens_render_bwd");

    variable_list out;
    out.push_back(need_ro ? g_ro : Tensor());
    out.push_back(need_rd ? g_rd : Tensor());
    for (int64_t k = 0; k < n_levels; ++k) {
      if (g_off[k] < 0) { out.push_back(Tensor()); continue; }
      Tensor gn = arena.narrow(0, g_off[k], nat[k].numel()).view(nat[k].sizes());           // [Z][Y][X][32]
      if (grid_native[k]) {
        out.push_back(gn.permute({3, 0, 1, 2}).unsqueeze(0));                               // a [1,32,Z,Y,X] view of it
      } else {
        Tensor g_ref = torch::empty({1, 32, gn.size(0), gn.size(1), gn.size(2)}, f32);
        ++g_launches;
        check(ens_grid_from_native((const float *)gn.data_ptr(), (float *)g_ref.data_ptr(), gn.size(0) * gn.size(1) * gn.size(2), 0,
                                   cur_stream(ro)), "ens_grid_from_native");
        out.push_back(g_ref);
      }
    }
    const auto shapes = ctx->saved_data["param_shapes"].toList();
    int64_t pi = 0;
    for (int64_t k = 0; k < n_levels; ++k) {
      int64_t off = d_off[k];
      for (int64_t j = 0; j < n_params[k]; ++j, ++pi) {
        const std::vector<int64_t> shp = shapes.get(pi).toIntVector();
        int64_t n = 1;
        for (int64_t d : shp) n *= d;
        if (any_param && ctx->needs_input_grad(2 + n_levels + pi)) out.push_back(arena.narrow(0, off, n).view(shp));
        else out.push_back(Tensor());
        off += n;
      }
    }
    // aux, native, packed: non-differentiable tensor inputs; then the scalar arguments
    const size_t n_tensor_tail = 4 + 2 * (size_t)n_levels;
    for (size_t i = 0; i < n_tensor_tail; ++i) out.push_back(Tensor());
    for (int i = 0; i < 9; ++i) out.push_back(Tensor());
    return out;
  }
};

std::vector<Tensor> render(const Tensor &rays_o, const Tensor &rays_d, std::vector<Tensor> grids, std::vector<Tensor> params,
                           std::vector<Tensor> aux, std::vector<Tensor> native, std::vector<Tensor> packed, int64_t stage,
                           int64_t n_samples, int64_t n_surface, std::vector<double> bounds, std::vector<int64_t> n_params,
                           bool want_aux, bool tc_map, bool tc_pose) {
  return RenderFn::apply(rays_o, rays_d, at::TensorList(grids), at::TensorList(params), at::TensorList(aux), at::TensorList(native),
                         at::TensorList(packed), stage, n_samples, n_surface, bounds, n_params, want_aux, at::GradMode::is_enabled(),
                         tc_map, tc_pose);
}

// ---------------------------------------------------------------------------------------------
// common.get_samples after the torch.randint draw (src/common.py:92-187); differentiable wrt c2w
// ---------------------------------------------------------------------------------------------
struct SampleRaysFn : public torch::autograd::Function<SampleRaysFn> {
  static variable_list forward(AutogradContext *ctx, const Tensor &c2w, const Tensor &indices, const Tensor &depth, const Tensor &color,
                               std::vector<int64_t> crop, std::vector<double> cam) {
    const c10::cuda::CUDAGuard guard(indices.device());
    const auto dev = indices.device();
    const int64_t n = indices.numel();
    const Tensor m = f32c(c2w);
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
    Tensor ro = torch::empty({n, 3}, f32), rd = torch::empty({n, 3}, f32), pi = torch::empty({n}, f32), pj = torch::empty({n}, f32);
    Tensor sd = torch::empty({n}, f32);
    TORCH_CHECK(color.scalar_type() == torch::kFloat32 || color.scalar_type() == torch::kFloat64, "color must be float32 or float64");
    const bool is64 = color.scalar_type() == torch::kFloat64;
    Tensor sc = torch::empty({n, 3}, color.options());
    Tensor dep = (depth.scalar_type() == torch::kFloat32 && depth.is_contiguous()) ? depth : depth.to(torch::kFloat32).contiguous();
    Tensor col = color.is_contiguous() ? color : color.contiguous();
    ++g_launches;
    check(ens_sample_rays((const int64_t *)indices.data_ptr(), n, (int)crop[0], (int)crop[1], (int)crop[2], (int)crop[3], (int)cam[0], (int)cam[1],
                          (float)cam[2], (float)cam[3], (float)cam[4], (float)cam[5], (const float *)m.data_ptr(), (int)m.stride(0),
                          (const float *)dep.data_ptr(), col.data_ptr(), is64 ? 1 : 0, (float *)pi.data_ptr(), (float *)pj.data_ptr(),
                          (float *)ro.data_ptr(), (float *)rd.data_ptr(), (float *)sd.data_ptr(), sc.data_ptr(), cur_stream(indices)),
          "ens_sample_rays");
    ctx->save_for_backward({pi, pj});
    ctx->saved_data["cam"] = cam;
    ctx->saved_data["rows"] = c2w.size(0);
    ctx->mark_non_differentiable({sd, sc});
    return {ro, rd, sd, sc};
  }
  static variable_list backward(AutogradContext *ctx, variable_list gout) {
    const variable_list sv = ctx->get_saved_variables();
    const Tensor &pi = sv[0], &pj = sv[1];
    const c10::cuda::CUDAGuard guard(pi.device());
    const std::vector<double> cam = ctx->saved_data["cam"].toDoubleVector();
    const int64_t rows = ctx->saved_data["rows"].toInt();
    Tensor g = torch::zeros({rows, 4}, pi.options());                  // a 4 x 4 c2w gets a zero last row
    Tensor gro, grd;
    if (gout[0].defined()) gro = f32c(gout[0]);
    if (gout[1].defined()) grd = f32c(gout[1]);
    ++g_launches;
    check(ens_rays_bwd((const float *)pi.data_ptr(), (const float *)pj.data_ptr(), pi.numel(), 0, (float)cam[2], (float)cam[3], (float)cam[4],
                       (float)cam[5], (const float *)ptr(gro), (const float *)ptr(grd), (float *)g.data_ptr(), cur_stream(pi)),
          "ens_rays_bwd");
    return {g, Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
  }
};

std::vector<Tensor> sample_rays(const Tensor &c2w, const Tensor &indices, const Tensor &depth, const Tensor &color,
                                std::vector<int64_t> crop, std::vector<double> cam) {
  return SampleRaysFn::apply(c2w, indices, depth, color, crop, cam);
}

// ---------------------------------------------------------------------------------------------
// common.get_camera_from_tensor / quad2rotation (src/common.py:189-228) for (n, 7) camera tensors
// ---------------------------------------------------------------------------------------------
struct PoseFn : public torch::autograd::Function<PoseFn> {
  static Tensor forward(AutogradContext *ctx, const Tensor &cam) {
    const c10::cuda::CUDAGuard guard(cam.device());
    const Tensor x = f32c(cam);
    const int64_t n = x.size(0);
    Tensor out = torch::empty({n, 3, 4}, x.options());
    ++g_launches;
    check(ens_pose_fwd((const float *)x.data_ptr(), (int)n, (float *)out.data_ptr(), cur_stream(x)), "ens_pose_fwd");
    ctx->save_for_backward({x});
    return out;
  }
  static variable_list backward(AutogradContext *ctx, variable_list gout) {
    const Tensor x = ctx->get_saved_variables()[0];
    const c10::cuda::CUDAGuard guard(x.device());
    const int64_t n = x.size(0);
    Tensor gc = torch::empty({n, 7}, x.options());
    const Tensor g = f32c(gout[0]);
    ++g_launches;
    check(ens_pose_bwd((const float *)x.data_ptr(), (int)n, (const float *)g.data_ptr(), (float *)gc.data_ptr(), cur_stream(x)), "ens_pose_bwd");
    return {gc};
  }
};

Tensor pose_to_c2w(const Tensor &cam) { return PoseFn::apply(cam); }

int64_t launches() { return g_launches; }
void reset_launches() { g_launches = 0; }

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("render", &render, "Renderer.render_batch_ray (fused forward + autograd backward)");
  m.def("sample_rays", &sample_rays, "get_samples after the pixel draw");
  m.def("pose_to_c2w", &pose_to_c2w, "get_camera_from_tensor for (n,7) camera tensors");
  m.def("launches", &launches);
  m.def("reset_launches", &reset_launches);
}
