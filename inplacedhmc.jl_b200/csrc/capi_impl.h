// capi_impl.h — the extern "C" entry points of include/bnuts.h, generated over an
// execution policy.  Included exactly once by a translation unit that defines
// BNUTS_EXEC (the policy type) before including this file.
#pragma once
#include "engine_core.h"

#ifndef BNUTS_EXEC
#error "define BNUTS_EXEC before including capi_impl.h"
#endif

namespace {
struct AnyEngine {
  int dtype;
  bn::EngineCore<double, BNUTS_EXEC>* e64 = nullptr;
  bn::EngineCore<float, BNUTS_EXEC>* e32 = nullptr;
};
thread_local std::string g_create_error;
}  // namespace

#define BN_DISPATCH(e, call)                                              \
  do {                                                                    \
    if (!(e)) return BNUTS_ERR_INVALID_ARGUMENT;                          \
    AnyEngine* ae = reinterpret_cast<AnyEngine*>(e);                      \
    try {                                                                 \
      if (ae->dtype == BNUTS_F64) { auto& E = *ae->e64; return E.call; }  \
      else { auto& E = *ae->e32; return E.call; }                         \
    } catch (const std::exception& ex) {                                  \
      if (ae->dtype == BNUTS_F64) ae->e64->err = ex.what(); else ae->e32->err = ex.what(); \
      return BNUTS_ERR_INTERNAL;                                          \
    }                                                                     \
  } while (0)

extern "C" {

int32_t bnuts_create(const bnuts_config* cfg, bnuts_engine** out) {
  if (!cfg || !out || cfg->n_chains <= 0 || cfg->dim <= 0 || cfg->max_depth <= 0 || cfg->max_depth > 32 ||
      !(cfg->min_delta < 0) || (cfg->dtype != BNUTS_F64 && cfg->dtype != BNUTS_F32)) {
    g_create_error = "invalid bnuts_config";
    return BNUTS_ERR_INVALID_ARGUMENT;
  }
  auto* ae = new AnyEngine();
  ae->dtype = cfg->dtype;
  int32_t rc;
  try {
    if (cfg->dtype == BNUTS_F64) { ae->e64 = new bn::EngineCore<double, BNUTS_EXEC>(); rc = ae->e64->init(*cfg); if (rc) g_create_error = ae->e64->err; }
    else { ae->e32 = new bn::EngineCore<float, BNUTS_EXEC>(); rc = ae->e32->init(*cfg); if (rc) g_create_error = ae->e32->err; }
  } catch (const std::exception& ex) { g_create_error = ex.what(); rc = BNUTS_ERR_INTERNAL; }
  if (rc) {   // release whatever init() had allocated before it failed (destroy() skips null pointers)
    if (ae->e64) { ae->e64->destroy(); delete ae->e64; }
    if (ae->e32) { ae->e32->destroy(); delete ae->e32; }
    delete ae;
    return rc;
  }
  *out = reinterpret_cast<bnuts_engine*>(ae);
  return 0;
}
int32_t bnuts_destroy(bnuts_engine* e) {
  if (!e) return 0;
  auto* ae = reinterpret_cast<AnyEngine*>(e);
  if (ae->e64) { ae->e64->destroy(); delete ae->e64; }
  if (ae->e32) { ae->e32->destroy(); delete ae->e32; }
  delete ae;
  return 0;
}
const char* bnuts_last_error(const bnuts_engine* e) {
  if (!e) return g_create_error.c_str();
  auto* ae = reinterpret_cast<const AnyEngine*>(e);
  return ae->dtype == BNUTS_F64 ? ae->e64->err.c_str() : ae->e32->err.c_str();
}
int32_t bnuts_model_iid_normal(bnuts_engine* e) { BN_DISPATCH(e, model_simple(bn::MODEL_IID_NORMAL)); }
int32_t bnuts_model_funnel(bnuts_engine* e) { BN_DISPATCH(e, model_simple(bn::MODEL_FUNNEL)); }
int32_t bnuts_model_gaussian(bnuts_engine* e, const double* P) { BN_DISPATCH(e, model_gaussian(P)); }
int32_t bnuts_model_logistic(bnuts_engine* e, const void* X, int32_t xd, const double* y, int64_t N, double tau,
                             int32_t rb) {
  BN_DISPATCH(e, model_logistic(X, xd, y, N, tau, rb));
}
int32_t bnuts_model_logistic_synthetic(bnuts_engine* e, uint64_t data_seed, int64_t row_offset, int64_t N, double tau, int32_t rb) {
  BN_DISPATCH(e, model_logistic_synth(data_seed, row_offset, N, tau, rb));
}
int32_t bnuts_synth_logistic_rows(uint64_t data_seed, int64_t row_offset, int64_t nrows, int32_t D, uint16_t* X, double* y,
                                  double* beta_true) {
  if (nrows < 0 || row_offset < 0 || D <= 0) return BNUTS_ERR_INVALID_ARGUMENT;
  bn::synth_rows_host(data_seed, row_offset, nrows, D, X, y, beta_true);
  return 0;
}
int32_t bnuts_logistic_set_reference(bnuts_engine* e, const double* beta_ref) { BN_DISPATCH(e, logistic_set_reference(beta_ref)); }
int32_t bnuts_set_positions(bnuts_engine* e, const double* q) { BN_DISPATCH(e, set_positions(q)); }
int32_t bnuts_get_state(bnuts_engine* e, double* q, double* g, double* l) { BN_DISPATCH(e, get_state(q, g, l)); }
int32_t bnuts_set_metric_diag(bnuts_engine* e, const double* m) { BN_DISPATCH(e, set_metric(m)); }
int32_t bnuts_get_metric_diag(bnuts_engine* e, double* m) {
  if (!m) return BNUTS_ERR_INVALID_ARGUMENT;
  BN_DISPATCH(e, get_metric(m));
}
int32_t bnuts_get_metric_diag_w(bnuts_engine* e, double* w) {
  if (!w) return BNUTS_ERR_INVALID_ARGUMENT;
  BN_DISPATCH(e, get_metric_w(w));
}
int32_t bnuts_set_metric_diag_pair(bnuts_engine* e, const double* m, const double* w) { BN_DISPATCH(e, set_metric_pair(m, w)); }
int32_t bnuts_set_metric_dense(bnuts_engine* e, const double* m) { BN_DISPATCH(e, set_metric_dense(m)); }
int32_t bnuts_get_metric_dense(bnuts_engine* e, double* m) {
  if (!m) return BNUTS_ERR_INVALID_ARGUMENT;
  BN_DISPATCH(e, get_metric_dense(m));
}
int32_t bnuts_set_stepsize(bnuts_engine* e, const double* eps) {
  if (!eps) return BNUTS_ERR_INVALID_ARGUMENT;
  BN_DISPATCH(e, set_stepsize(eps));
}
int32_t bnuts_get_stepsize(bnuts_engine* e, double* eps) {
  if (!eps) return BNUTS_ERR_INVALID_ARGUMENT;
  BN_DISPATCH(e, get_stepsize(eps));
}
int32_t bnuts_seed(bnuts_engine* e, uint64_t seed, uint32_t next_t) {
  if (!e) return BNUTS_ERR_INVALID_ARGUMENT;
  AnyEngine* ae = reinterpret_cast<AnyEngine*>(e);
  if (ae->dtype == BNUTS_F64) { ae->e64->rp.seed = seed; ae->e64->next_t = next_t; }
  else { ae->e32->rp.seed = seed; ae->e32->next_t = next_t; }
  return 0;
}
int32_t bnuts_get_rng(bnuts_engine* e, uint64_t* seed, uint32_t* next_t) {
  if (!e || !seed || !next_t) return BNUTS_ERR_INVALID_ARGUMENT;
  AnyEngine* ae = reinterpret_cast<AnyEngine*>(e);
  if (ae->dtype == BNUTS_F64) { *seed = ae->e64->rp.seed; *next_t = ae->e64->next_t; }
  else { *seed = ae->e32->rp.seed; *next_t = ae->e32->next_t; }
  return 0;
}
int32_t bnuts_inject(bnuts_engine* e, int32_t T, const uint32_t* dirs, const double* p, const double* exps, int32_t n_exps) {
  BN_DISPATCH(e, inject(T, dirs, p, exps, n_exps));
}
int32_t bnuts_leapfrog(bnuts_engine* e, const double* p_in, const double* eps, int32_t nsteps, double* q_out,
                       double* p_out, double* g_out, double* l_out) {
  BN_DISPATCH(e, leapfrog(p_in, eps, nsteps, q_out, p_out, g_out, l_out));
}
int32_t bnuts_find_local_optimum(bnuts_engine* e, double magnitude_penalty, int32_t iterations) {
  BN_DISPATCH(e, find_local_optimum(magnitude_penalty, iterations));
}
int32_t bnuts_find_initial_stepsize(bnuts_engine* e, const bnuts_stepsize_search* P) {
  if (!P) return BNUTS_ERR_INVALID_ARGUMENT;
  BN_DISPATCH(e, find_initial_stepsize(*P));
}
int32_t bnuts_warmup_stage(bnuts_engine* e, int32_t N, int32_t metric_kind, const bnuts_dual_averaging* da,
                           double lambda, double* chain_out, int64_t sd, int64_t sc, bnuts_tree_stats* stats_out,
                           int64_t ssc, double* eps_out) {
  // da == NULL ≙ FixedStepsize (src/stepsize.jl:251-255): the stage keeps ϵ and only tunes the metric
  BN_DISPATCH(e, transitions(N, da, metric_kind, lambda, chain_out, sd, sc, stats_out, ssc, nullptr, eps_out));
}
int32_t bnuts_sample(bnuts_engine* e, int32_t N, double* chain_out, int64_t sd, int64_t sc,
                     bnuts_tree_stats* stats_out, int64_t ssc, int32_t* sel) {
  BN_DISPATCH(e, transitions(N, nullptr, BNUTS_METRIC_NONE, 0.0, chain_out, sd, sc, stats_out, ssc, sel, nullptr));
}
int32_t bnuts_set_allreduce(bnuts_engine* e, bnuts_allreduce_fn fn, void* ctx) {
  if (!fn) return BNUTS_ERR_INVALID_ARGUMENT;
  BN_DISPATCH(e, enable_reduce(fn, ctx));
}
int32_t bnuts_nccl_unique_id(uint8_t* id) { return id ? BNUTS_EXEC::nccl_unique_id(id) : BNUTS_ERR_INVALID_ARGUMENT; }
int32_t bnuts_set_nccl(bnuts_engine* e, const uint8_t* id, int32_t world, int32_t rank) {
  if (!e || !id || world < 1 || rank < 0 || rank >= world) return BNUTS_ERR_INVALID_ARGUMENT;
  AnyEngine* ae = reinterpret_cast<AnyEngine*>(e);
  std::string& err = ae->dtype == BNUTS_F64 ? ae->e64->err : ae->e32->err;
  int32_t rc = ae->dtype == BNUTS_F64 ? ae->e64->x.nccl_init(id, world, rank, err) : ae->e32->x.nccl_init(id, world, rank, err);
  if (rc) return rc;
  if (ae->dtype == BNUTS_F64) ae->e64->red_world = world; else ae->e32->red_world = world;
  BN_DISPATCH(e, enable_reduce(nullptr, nullptr));
}
int32_t bnuts_p2p_export(bnuts_engine* e, uint8_t* handle) {
  if (!handle) return BNUTS_ERR_INVALID_ARGUMENT;
  BN_DISPATCH(e, p2p_export(handle));
}
int32_t bnuts_p2p_connect(bnuts_engine* e, const uint8_t* handles, int32_t world, int32_t rank) {
  if (!handles || world < 1 || world > 8 || rank < 0 || rank >= world) return BNUTS_ERR_INVALID_ARGUMENT;
  BN_DISPATCH(e, p2p_connect(handles, world, rank));
}
int32_t bnuts_counters(bnuts_engine* e, bnuts_counter_block* out) {
  if (!out) return BNUTS_ERR_INVALID_ARGUMENT;
  BN_DISPATCH(e, get_counters(out));
}
int32_t bnuts_profile(bnuts_engine* e, int32_t enable, double* ms, int64_t* launches) {
  BN_DISPATCH(e, x.profile(enable, ms, launches));
}
int32_t bnuts_chain_status(bnuts_engine* e, int32_t* st) {
  if (!st) return BNUTS_ERR_INVALID_ARGUMENT;
  BN_DISPATCH(e, chain_status(st));
}

}  // extern "C"
