// gecut_api.cu -- the C ABI (include/gecut.h): contexts, orchestration of the kernels, host<->device copies.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "gecut_internal.cuh"

// ------------------------------------------------------------------------------------------------
// memory: stream-ordered pool allocations so that repeated cuts reuse memory without cudaMalloc round trips
// ------------------------------------------------------------------------------------------------
void gc_trim(gecut_ctx* ctx) {
  if (ctx->cache.empty()) return;
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->cache) cudaFree(kv.second);
  ctx->cache.clear();
  ctx->cached_bytes = 0;
}

int gc_malloc(gecut_ctx* ctx, void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  bytes = (bytes + 255) & ~(size_t)255;
  // smallest cached block that fits without wasting more than 1/8 of it
  auto it = ctx->cache.lower_bound(bytes);
  if (it != ctx->cache.end() && it->first <= bytes + bytes / 8 + 4096) {
    *p = it->second;
    ctx->cached_bytes -= it->first;
    ctx->live[*p] = it->first;
    ctx->cache.erase(it);
    return GECUT_OK;
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {  // give the cached blocks back and retry once
    (void)cudaGetLastError();
    gc_trim(ctx);
    e = cudaMalloc(p, bytes);
  }
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    char buf[128];
    snprintf(buf, sizeof(buf), "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    ctx->err = buf;
    *p = nullptr;
    return e == cudaErrorMemoryAllocation ? GECUT_ERR_OOM : GECUT_ERR_CUDA;
  }
  ctx->live[*p] = bytes;
  return GECUT_OK;
}

void gc_free(gecut_ctx* ctx, void* p) {
  if (!p) return;
  auto it = ctx->live.find(p);
  if (it == ctx->live.end()) return;  // not ours (or already freed)
  ctx->cache.insert({it->second, p});
  ctx->cached_bytes += it->second;
  ctx->live.erase(it);
}

int gc_launch_check(gecut_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
    return GECUT_ERR_CUDA;
  }
  return GECUT_OK;
}

static inline int64_t pad_pitch(int64_t n) { return ((n + 255) / 256) * 256; }

void gc_prof_before(gecut_ctx* ctx, const char* name) {
  GcProfRec r;
  r.name = name;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, ctx->stream);
  ctx->prof.push_back(r);
}

void gc_prof_after(gecut_ctx* ctx) { cudaEventRecord(ctx->prof.back().b, ctx->stream); }

namespace {

int make_grid(gecut_ctx* ctx, const gecut_grid* in, GcGrid* g) {
  if (!in) {
    ctx->err = "grid is NULL";
    return GECUT_ERR_INVALID;
  }
  if (in->D != 2 && in->D != 3) {
    ctx->err = "only 2-D and 3-D background models are implemented";
    return GECUT_ERR_NOTIMPLEMENTED;
  }
  memset(g, 0, sizeof(*g));
  const int D = in->D;
  g->D = D;
  g->simplexified = in->simplexified ? 1 : 0;
  for (int d = 0; d < 3; ++d) {
    if (d < D) {
      if (in->partition[d] <= 0 || !(in->pmax[d] > in->pmin[d])) {
        ctx->err = "invalid Cartesian descriptor (partition <= 0 or pmax <= pmin)";
        return GECUT_ERR_INVALID;
      }
      g->n[d] = in->partition[d];
      g->nn[d] = in->partition[d] + 1;
      g->pmin[d] = in->pmin[d];
      g->dx[d] = (in->pmax[d] - in->pmin[d]) / (double)in->partition[d];  // CartesianDescriptor sizes
    } else {
      g->n[d] = 1;
      g->nn[d] = 1;
      g->pmin[d] = 0.0;
      g->dx[d] = 0.0;
    }
  }
  g->k0 = 0;
  g->k1 = g->n[D - 1];
  if (!(in->slab_begin == 0 && in->slab_end == 0)) {
    if (in->slab_begin < 0 || in->slab_end > g->n[D - 1] || in->slab_begin >= in->slab_end) {
      ctx->err = "invalid slab range";
      return GECUT_ERR_INVALID;
    }
    g->k0 = in->slab_begin;
    g->k1 = in->slab_end;
  }
  g->nt = D == 2 ? 2 : 6;
  g->cpc = g->simplexified ? g->nt : 1;
  g->nt0 = g->simplexified ? 1 : g->nt;
  g->nv = g->simplexified ? D + 1 : (1 << D);
  g->layer_cubes = D == 3 ? (int64_t)g->n[0] * g->n[1] : (int64_t)g->n[0];
  g->layer_nodes = D == 3 ? (int64_t)g->nn[0] * g->nn[1] : (int64_t)g->nn[0];
  g->num_cubes = g->layer_cubes * (g->k1 - g->k0);
  g->num_cells = g->num_cubes * g->cpc;
  g->num_nodes = g->layer_nodes * (g->k1 - g->k0 + 1);
  if (g->num_cells >= 2147483647ll || g->layer_nodes * (int64_t)g->nn[D - 1] >= 2147483647ll) {
    ctx->err = "background model too large for the Int32 ids of the reference layout";
    return GECUT_ERR_OVERFLOW;
  }
  return GECUT_OK;
}

void free_result_arrays(gecut_result* r) {
  gecut_ctx* ctx = r->ctx;
  GcLayout& l = r->lay;
  gc_free(ctx, l.bg_state);
  gc_free(ctx, l.cutcells);
  gc_free(ctx, l.sc_c2p);
  gc_free(ctx, l.sc_c2bg);
  gc_free(ctx, l.sc_X);
  gc_free(ctx, l.sc_R);
  gc_free(ctx, l.sc_state);
  gc_free(ctx, l.sf_c2p);
  gc_free(ctx, l.sf_c2bg);
  gc_free(ctx, l.sf_N);
  if (!r->sf_points_alias) {
    gc_free(ctx, l.sf_X);
    gc_free(ctx, l.sf_R);
  }
  gc_free(ctx, l.sf_state);
  gc_free(ctx, l.sc_meas);
  gc_free(ctx, l.sf_meas);
  gc_free(ctx, r->cutmask);
  gc_free(ctx, r->cutcount);
  gc_free(ctx, r->cut22mask);
  gc_free(ctx, r->rowcount);
  r->rowcount = nullptr;
  gc_free(ctx, r->tile_n22);
  r->cut22mask = nullptr;
  r->tile_n22 = nullptr;
  for (auto& kv : r->prog_rows) {
    gc_free(ctx, kv.second.first);
    gc_free(ctx, kv.second.second);
  }
  r->prog_rows.clear();
  gc_free(ctx, r->cut_sc_ptr);
  r->cut_sc_ptr = nullptr;
  gc_free(ctx, r->cell_vio);
  r->cell_vio = nullptr;
  gc_free(ctx, r->bg_counts_dev);
  r->bg_counts_dev = nullptr;
  for (int i = 0; i < GC_LMAX; ++i)
    if (r->owns_nodal[i]) gc_free(ctx, r->nodal_dev[i]);
  l = GcLayout{};
  r->cutmask = nullptr;
  r->cutcount = nullptr;
}

int do_cut(gecut_ctx* ctx, const gecut_grid* grid, const gecut_leaf* leaves, int L, const double* const* nodal_values,
           int nodal_on_device, bool classify_only, gecut_result** out) {
  if (!ctx) return GECUT_ERR_INVALID;
  if (!out || !leaves) {
    ctx->err = "NULL argument";
    return GECUT_ERR_INVALID;
  }
  *out = nullptr;
  if (L < 1 || L > GC_LMAX) {
    ctx->err = "number of level sets must be in 1..GECUT_MAX_LEVELSETS";
    return L < 1 ? GECUT_ERR_INVALID : GECUT_ERR_NOTIMPLEMENTED;
  }
  DeviceGuard guard(ctx->device);
  gecut_result* r = new (std::nothrow) gecut_result();
  if (!r) return GECUT_ERR_OOM;
  r->ctx = ctx;
  r->L = L;
  int st = make_grid(ctx, grid, &r->grid);
  if (st != GECUT_OK) {
    delete r;
    return st;
  }
  const GcGrid& g = r->grid;
  auto fail = [&](int code) {
    free_result_arrays(r);
    cudaStreamSynchronize(ctx->stream);
    delete r;
    return code;
  };
  // leaves
  r->leaves.L = L;
  const bool nodal_local = grid->nodal_slab_local != 0;
  r->leaves.nodal_base = nodal_local ? (int64_t)g.k0 * g.layer_nodes : 0;
  const int64_t nn_total = nodal_local ? g.num_nodes : g.layer_nodes * g.nn[g.D - 1];
  for (int i = 0; i < L; ++i) {
    r->leaves.leaf[i] = leaves[i];
    r->leaves.nodal[i] = nullptr;
    const int kind = leaves[i].kind;
    if (kind < GECUT_LEAF_SPHERE || kind > GECUT_LEAF_NODAL) {
      ctx->err = "unknown leaf kind";
      return fail(GECUT_ERR_INVALID);
    }
    if (kind == GECUT_LEAF_DOUGHNUT && g.D != 3) {
      ctx->err = "doughnut leaves need a 3-D model";
      return fail(GECUT_ERR_INVALID);
    }
    if (kind == GECUT_LEAF_NODAL) {
      if (!nodal_values || !nodal_values[i]) {
        ctx->err = "NODAL leaf without nodal values";
        return fail(GECUT_ERR_INVALID);
      }
      if (nodal_on_device) {
        r->leaves.nodal[i] = nodal_values[i];
      } else {
        double* d = nullptr;
        st = gc_malloc(ctx, (void**)&d, sizeof(double) * nn_total);
        if (st != GECUT_OK) return fail(st);
        r->nodal_dev[i] = d;
        r->owns_nodal[i] = true;
        cudaError_t e = cudaMemcpyAsync(d, nodal_values[i], sizeof(double) * nn_total, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) {
          ctx->err = std::string("H2D copy of nodal values: ") + cudaGetErrorString(e);
          return fail(GECUT_ERR_CUDA);
        }
        r->leaves.nodal[i] = d;
      }
    }
  }
  // sizes
  gecut_sizes& sz = r->sizes;
  memset(&sz, 0, sizeof(sz));
  sz.num_bgcells = g.num_cells;
  sz.num_bgnodes = g.num_nodes;
  sz.cell_offset = (int64_t)g.k0 * g.layer_cubes * g.cpc;
  sz.node_offset = (int64_t)g.k0 * g.layer_nodes;
  sz.num_levelsets = L;
  sz.D = g.D;
  sz.num_vertices_per_bgcell = g.nv;
  // classification
  r->lay.bg_pitch = pad_pitch(g.num_cells);
  st = gc_malloc(ctx, (void**)&r->lay.bg_state, (size_t)r->lay.bg_pitch * L);
  if (st != GECUT_OK) return fail(st);
  const int wxtot = (g.n[0] + 31) / 32;
  const int64_t rows = (g.D == 3) ? (int64_t)(g.k1 - g.k0) * g.n[1] : (int64_t)(g.k1 - g.k0);
  const size_t mask_bytes = sizeof(uint32_t) * rows * wxtot * g.cpc;
  st = gc_malloc(ctx, (void**)&r->cutmask, mask_bytes);
  if (st != GECUT_OK) return fail(st);
  st = gc_malloc(ctx, (void**)&r->bg_counts_dev, sizeof(unsigned long long) * GC_LMAX * 12);
  if (st != GECUT_OK) return fail(st);
  if (cudaMemsetAsync(r->bg_counts_dev, 0, sizeof(unsigned long long) * GC_LMAX * 12, ctx->stream) != cudaSuccess) {
    ctx->err = "memset of the bg counters failed";
    return fail(GECUT_ERR_CUDA);
  }
  r->want_n22 = !classify_only && L == 1 && g.simplexified && g.D == 3;
  st = gc_classify(ctx, r);
  if (st != GECUT_OK) return fail(st);
  st = gc_compact_cutcells(ctx, r);
  if (st != GECUT_OK) return fail(st);
  if (!classify_only) {
    st = gc_cut_cells(ctx, r);
    if (st != GECUT_OK) return fail(st);
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    ctx->err = std::string("cut: ") + cudaGetErrorString(e);
    return fail(GECUT_ERR_CUDA);
  }
  *out = r;
  return GECUT_OK;
}

int make_program(gecut_ctx* ctx, const int32_t* program, int len, int L, GcProgram* p) {
  if (!program || len <= 0 || len > GECUT_MAX_PROGRAM) {
    ctx->err = "invalid boolean program";
    return GECUT_ERR_INVALID;
  }
  int depth = 0;
  for (int i = 0; i < len; ++i) {
    int t = program[i];
    if (t >= 0) {
      if (t >= L) {
        ctx->err = "program refers to a level set that does not exist";
        return GECUT_ERR_INVALID;
      }
      depth++;
    } else if (t == GECUT_OP_COMPLEMENT) {
      if (depth < 1) goto bad;
    } else if (t >= GECUT_OP_SETDIFF) {
      if (depth < 2) goto bad;
      depth--;
    } else {
      goto bad;
    }
    p->tok[i] = (int8_t)t;
  }
  if (depth != 1) goto bad;
  p->len = len;
  p->lut = nullptr;
  p->blut = nullptr;
  p->nused = 0;
  p->sc_base = p->sc_row = p->bg_base = p->bg_row = nullptr;
  if (len >= 3) {
    // truth table over the level sets the program uses (EmbeddedDiscretizations.jl:136-177 evaluated 3^nused times, once)
    bool seen[GC_LMAX] = {false};
    for (int i = 0; i < len; ++i)
      if (program[i] >= 0) seen[program[i]] = true;
    int nused = 0;
    for (int ls = 0; ls < L; ++ls)
      if (seen[ls]) {
        if (nused < GC_LUT_MAXUSED) p->used[nused] = (int8_t)ls;
        nused++;
      }
    if (nused <= GC_LUT_MAXUSED) {
      p->nused = nused;
      const std::string key(reinterpret_cast<const char*>(p->tok), (size_t)len);
      auto it = ctx->lut_cache.find(key);
      if (it == ctx->lut_cache.end()) {
        size_t n = 1;
        for (int k = 0; k < nused; ++k) n *= 3;
        std::vector<int8_t> tab(n);
        int8_t vals[GC_LMAX] = {0};
        for (size_t idx = 0; idx < n; ++idx) {
          size_t r = idx;
          for (int k = 0; k < nused; ++k) {
            vals[p->used[k]] = (int8_t)((int)(r % 3) - 1);
            r /= 3;
          }
          tab[idx] = (int8_t)gc_eval_inoutcut_vals(*p, vals);
        }
        int8_t* d = nullptr;
        if (cudaMalloc((void**)&d, n) != cudaSuccess) {
          ctx->err = "out of device memory (truth table)";
          return GECUT_ERR_OOM;
        }
        it = ctx->lut_cache.emplace(key, std::make_pair(d, std::move(tab))).first;
        if (cudaMemcpyAsync(d, it->second.second.data(), n, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
          ctx->err = "upload of the truth table failed";
          return GECUT_ERR_CUDA;
        }
      }
      p->lut = it->second.first;
    }
  }
  return GECUT_OK;
bad:
  ctx->err = "malformed boolean program";
  return GECUT_ERR_INVALID;
}

// compute_inoutboundary (EmbeddedDiscretizations.jl:179-245) as a truth table over the level sets the program uses: entry =
// (state + 1) | (orientation + 1) << 2, same base-3 index as `lut` (subfacet states are IN / OUT / INTERFACE = -1 / 1 / 0)
static int attach_boundary_lut(gecut_ctx* ctx, GcProgram* p) {
  if (p->len < 3 || p->nused <= 0 || p->nused > GC_LUT_MAXUSED) return GECUT_OK;  // single leaf: no table needed
  const std::string key = "B" + std::string(reinterpret_cast<const char*>(p->tok), (size_t)p->len);
  auto it = ctx->lut_cache.find(key);
  if (it == ctx->lut_cache.end()) {
    size_t n = 1;
    for (int k = 0; k < p->nused; ++k) n *= 3;
    std::vector<int8_t> tab(n);
    int8_t vals[GC_LMAX] = {0};
    for (size_t idx = 0; idx < n; ++idx) {
      size_t r = idx;
      for (int k = 0; k < p->nused; ++k) {
        vals[p->used[k]] = (int8_t)((int)(r % 3) - 1);
        r /= 3;
      }
      int st = 0, orient = 1;
      gc_eval_inoutboundary_vals(*p, vals, &st, &orient);
      tab[idx] = (int8_t)((st + 1) | ((orient + 1) << 2));
    }
    int8_t* d = nullptr;
    if (cudaMalloc((void**)&d, n) != cudaSuccess) {
      ctx->err = "out of device memory (boundary truth table)";
      return GECUT_ERR_OOM;
    }
    it = ctx->lut_cache.emplace(key, std::make_pair(d, std::move(tab))).first;
    if (cudaMemcpyAsync(d, it->second.second.data(), n, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
      ctx->err = "upload of the boundary truth table failed";
      return GECUT_ERR_CUDA;
    }
  }
  p->blut = it->second.first;
  return GECUT_OK;
}

// Multi-operator programs: evaluate once over the subcell and bg-cell rows of `res`, keep the two Int8 rows with the
// result and let every later evaluation of the same program read them (selections of both sides, measures, Laplacian).
int attach_program_rows(gecut_ctx* ctx, const gecut_result* res_c, GcProgram* p) {
  // single-leaf programs read one state row anyway.  Multi-operator programs: the SUBCELL row is always memoised (O(Nsc)
  // bytes, one short pass; every later consumer -- both sides' selections, the Laplacian's per-subcell test -- reads one byte
  // instead of the dependent chain state rows -> truth-table gather); the BG-CELL row (a pass over all Nc cells) only when
  // the program uses more than four level sets, else the truth-table gather over the rows is cheaper than writing a new row
  if (!p->lut) return GECUT_OK;
  const bool want_bg = p->nused > 4;
  gecut_result* res = const_cast<gecut_result*>(res_c);
  const std::string key(reinterpret_cast<const char*>(p->tok), (size_t)p->len);
  auto it = res->prog_rows.find(key);
  if (it == res->prog_rows.end()) {
    const int64_t nsc = res->sizes.num_subcells, nc = res->sizes.num_bgcells;
    int8_t *sc = nullptr, *bg = nullptr;
    if (nsc > 0) {
      GC_CHECK(gc_malloc(ctx, (void**)&sc, (size_t)pad_pitch(nsc)));  // padded: consumers read 16-byte vectors
      GC_CHECK(gc_eval_rows(ctx, res->lay.sc_state, res->lay.sc_pitch, nsc, *p, sc, true));
    }
    if (nc > 0 && want_bg) {
      GC_CHECK(gc_malloc(ctx, (void**)&bg, (size_t)pad_pitch(nc)));
      GC_CHECK(gc_eval_rows(ctx, res->lay.bg_state, res->lay.bg_pitch, nc, *p, bg, true));
    }
    it = res->prog_rows.emplace(key, std::make_pair(sc, bg)).first;
  }
  p->sc_base = res->lay.sc_state;
  p->sc_row = it->second.first;
  p->bg_base = res->lay.bg_state;
  p->bg_row = it->second.second;
  return GECUT_OK;
}

int copy_rows_to_host(gecut_ctx* ctx, int8_t* host, const int8_t* dev, int64_t pitch, int64_t n, int L) {
  if (!host || n == 0) return GECUT_OK;
  GC_CUDA(cudaMemcpy2DAsync(host, (size_t)n, dev, (size_t)pitch, (size_t)n, (size_t)L, cudaMemcpyDeviceToHost, ctx->stream));
  return GECUT_OK;
}

template <class T>
int copy_to_host(gecut_ctx* ctx, T* host, const T* dev, int64_t count) {
  if (!host || count == 0 || !dev) return GECUT_OK;
  GC_CUDA(cudaMemcpyAsync(host, dev, sizeof(T) * count, cudaMemcpyDeviceToHost, ctx->stream));
  return GECUT_OK;
}

}  // namespace

// shared with gecut_facets.cu
int gc_make_grid(gecut_ctx* ctx, const gecut_grid* in, GcGrid* g) { return make_grid(ctx, in, g); }
int gc_make_program(gecut_ctx* ctx, const int32_t* program, int len, int L, GcProgram* p) {
  return make_program(ctx, program, len, L, p);
}

// ------------------------------------------------------------------------------------------------
// ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int gecut_create(int device, void* cuda_stream, gecut_ctx** out) {
  if (!out) return GECUT_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    (void)cudaGetLastError();
    return GECUT_ERR_CUDA;  // no CUDA device: there is no CPU fallback
  }
  gecut_ctx* ctx = new (std::nothrow) gecut_ctx();
  if (!ctx) return GECUT_ERR_OOM;
  ctx->device = device;
  ctx->stream = (cudaStream_t)cuda_stream;
  DeviceGuard guard(device);
  // keep freed blocks in the pool: repeated cuts then reuse memory without touching the driver
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  // behind the three tables: packed vertex order -> packed order after an all-IN / all-OUT stage (cm_skip_*_stages)
  uint8_t perm_lut[3 * 512];
  for (int di = 0; di < 3; ++di) {
    const int n = di + 2;  // vertices of the simplex
    for (int b = 0; b < 2; ++b) {
      const uint8_t* pm = GC_SIMPLEX_TABLES[di].sub_pts[b ? (1 << n) - 1 : 0][0];
      for (int q = 0; q < 256; ++q) {
        int out = 0;
        for (int i = 0; i < n; ++i) out |= ((q >> (2 * pm[i])) & 3) << (2 * i);
        perm_lut[di * 512 + b * 256 + q] = (uint8_t)out;
      }
    }
  }
  bool ok = cudaMalloc((void**)&ctx->d_tables, sizeof(GC_SIMPLEX_TABLES) + sizeof(perm_lut)) == cudaSuccess &&
            cudaMemcpy(reinterpret_cast<uint8_t*>(ctx->d_tables) + sizeof(GC_SIMPLEX_TABLES), perm_lut, sizeof(perm_lut),
                       cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMalloc((void**)&ctx->d_simplexify_raw, sizeof(GC_SIMPLEXIFY_RAW)) == cudaSuccess &&
            cudaMalloc((void**)&ctx->d_simplexify_pos, sizeof(GC_SIMPLEXIFY_POS)) == cudaSuccess &&
            cudaMemcpy(ctx->d_tables, GC_SIMPLEX_TABLES, sizeof(GC_SIMPLEX_TABLES), cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(ctx->d_simplexify_raw, GC_SIMPLEXIFY_RAW, sizeof(GC_SIMPLEXIFY_RAW), cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(ctx->d_simplexify_pos, GC_SIMPLEXIFY_POS, sizeof(GC_SIMPLEXIFY_POS), cudaMemcpyHostToDevice) == cudaSuccess;
  ctx->h_pinned_bytes = 4096;
  ok = ok && cudaMallocHost(&ctx->h_pinned, ctx->h_pinned_bytes) == cudaSuccess;
  if (!ok) {
    (void)cudaGetLastError();
    gecut_destroy(ctx);
    return GECUT_ERR_CUDA;
  }
  *out = ctx;
  return GECUT_OK;
}

int gecut_destroy(gecut_ctx* ctx) {
  if (!ctx) return GECUT_OK;
  DeviceGuard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  gecut_comm_free(ctx);
  gc_trim(ctx);
  for (auto& kv : ctx->live) cudaFree(kv.first);
  ctx->live.clear();
  cudaFree(ctx->d_tables);
  cudaFree(ctx->d_simplexify_raw);
  cudaFree(ctx->d_simplexify_pos);
  for (auto& kv : ctx->lut_cache) cudaFree(kv.second.first);
  for (int i = 0; i < 8; ++i) {
    if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
    if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  delete ctx;
  return GECUT_OK;
}

int gecut_set_stream(gecut_ctx* ctx, void* cuda_stream) {
  if (!ctx) return GECUT_ERR_INVALID;
  DeviceGuard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  return GECUT_OK;
}

const char* gecut_last_error(const gecut_ctx* ctx) { return ctx ? ctx->err.c_str() : "NULL context"; }

int64_t gecut_launch_count(const gecut_ctx* ctx) { return ctx ? ctx->launches : 0; }

int gecut_host_alloc(size_t bytes, void** ptr) {
  if (!ptr) return GECUT_ERR_INVALID;
  *ptr = nullptr;
  if (bytes == 0) return GECUT_OK;
  const cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *ptr = nullptr;
    return e == cudaErrorMemoryAllocation ? GECUT_ERR_OOM : GECUT_ERR_CUDA;
  }
  return GECUT_OK;
}

int gecut_host_free(void* ptr) {
  if (!ptr) return GECUT_OK;
  return cudaFreeHost(ptr) == cudaSuccess ? GECUT_OK : GECUT_ERR_CUDA;
}

int gecut_host_register(void* ptr, size_t bytes) {
  if (!ptr || bytes == 0) return GECUT_ERR_INVALID;
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? GECUT_ERR_OOM : GECUT_ERR_CUDA;
  }
  return GECUT_OK;
}

int gecut_host_unregister(void* ptr) {
  if (!ptr) return GECUT_ERR_INVALID;
  return cudaHostUnregister(ptr) == cudaSuccess ? GECUT_OK : GECUT_ERR_CUDA;
}

int gecut_trim(gecut_ctx* ctx) {
  if (!ctx) return GECUT_ERR_INVALID;
  DeviceGuard guard(ctx->device);
  gc_trim(ctx);
  return GECUT_OK;
}

int gecut_profile(gecut_ctx* ctx, int enable) {
  if (!ctx) return GECUT_ERR_INVALID;
  DeviceGuard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& r : ctx->prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  ctx->prof.clear();
  ctx->prof_on = enable != 0;
  return GECUT_OK;
}

const char* gecut_profile_report(gecut_ctx* ctx) {
  if (!ctx) return "";
  DeviceGuard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  // aggregate by kernel name, in order of first launch: "name count total_ms\n"
  std::vector<std::pair<std::string, std::pair<int, double>>> agg;
  for (auto& r : ctx->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    bool found = false;
    for (auto& a : agg)
      if (a.first == r.name) {
        a.second.first++;
        a.second.second += ms;
        found = true;
        break;
      }
    if (!found) agg.push_back({r.name, {1, (double)ms}});
  }
  ctx->prof_report.clear();
  char buf[256];
  for (auto& a : agg) {
    snprintf(buf, sizeof(buf), "%s %d %.6f\n", a.first.c_str(), a.second.first, a.second.second);
    ctx->prof_report += buf;
  }
  return ctx->prof_report.c_str();
}

int gecut_cut(gecut_ctx* ctx, const gecut_grid* grid, const gecut_leaf* leaves, int num_leaves,
              const double* const* nodal_values, int nodal_on_device, gecut_result** out) {
  return do_cut(ctx, grid, leaves, num_leaves, nodal_values, nodal_on_device, false, out);
}

int gecut_classify(gecut_ctx* ctx, const gecut_grid* grid, const gecut_leaf* leaves, int num_leaves,
                   const double* const* nodal_values, int nodal_on_device, gecut_result** out) {
  return do_cut(ctx, grid, leaves, num_leaves, nodal_values, nodal_on_device, true, out);
}

int gecut_eval_leaf_at_nodes(gecut_ctx* ctx, const gecut_grid* grid, const gecut_leaf* leaf, double* values,
                             int on_device) {
  if (!ctx) return GECUT_ERR_INVALID;
  if (!grid || !leaf || !values) {
    ctx->err = "NULL argument";
    return GECUT_ERR_INVALID;
  }
  if (leaf->kind < GECUT_LEAF_SPHERE || leaf->kind >= GECUT_LEAF_NODAL) {
    ctx->err = "only closed-form leaves can be evaluated on the device";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  GcGrid g;
  gecut_grid gg = *grid;
  // nodal_slab_local: only the node planes slab_begin..slab_end (what a rank of the z-slab partition owns + its ghost plane)
  const bool slab_only = gg.nodal_slab_local != 0 && !(gg.slab_begin == 0 && gg.slab_end == 0);
  if (!slab_only) gg.slab_begin = gg.slab_end = 0;
  GC_CHECK(make_grid(ctx, &gg, &g));
  const int64_t nn = slab_only ? g.num_nodes : g.layer_nodes * g.nn[g.D - 1];
  double* d = values;
  if (!on_device) GC_CHECK(gc_malloc(ctx, (void**)&d, sizeof(double) * nn));
  int st = gc_eval_leaf_nodes(ctx, g, *leaf, d, slab_only);
  if (st == GECUT_OK && !on_device) {
    cudaError_t e = cudaMemcpyAsync(values, d, sizeof(double) * nn, cudaMemcpyDeviceToHost, ctx->stream);
    if (e != cudaSuccess) {
      ctx->err = cudaGetErrorString(e);
      st = GECUT_ERR_CUDA;
    }
  }
  if (!on_device) gc_free(ctx, d);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (st == GECUT_OK && e != cudaSuccess) {
    ctx->err = cudaGetErrorString(e);
    st = GECUT_ERR_CUDA;
  }
  return st;
}

int gecut_result_free(gecut_result* res) {
  if (!res) return GECUT_OK;
  if (res->live_selections > 0) {  // deferred: the last selection of this result releases it
    res->free_requested = true;
    return GECUT_OK;
  }
  DeviceGuard guard(res->ctx->device);
  free_result_arrays(res);
  delete res;
  return GECUT_OK;
}

int gecut_result_sizes(const gecut_result* res, gecut_sizes* sizes) {
  if (!res || !sizes) return GECUT_ERR_INVALID;
  *sizes = res->sizes;
  return GECUT_OK;
}

int gecut_result_device_ptrs(const gecut_result* res, gecut_buffers* dev) {
  if (!res || !dev) return GECUT_ERR_INVALID;
  const GcLayout& l = res->lay;
  dev->ls_to_bgcell_to_inoutcut = l.bg_state;
  dev->cutcell_to_cell = l.cutcells;
  dev->sc_cell_to_points = l.sc_c2p;
  dev->sc_cell_to_bgcell = l.sc_c2bg;
  dev->sc_point_to_coords = l.sc_X;
  dev->sc_point_to_rcoords = l.sc_R;
  dev->ls_to_subcell_to_inout = l.sc_state;
  dev->sf_facet_to_points = l.sf_c2p;
  dev->sf_facet_to_bgcell = l.sf_c2bg;
  dev->sf_facet_to_normal = l.sf_N;
  dev->sf_point_to_coords = l.sf_X;
  dev->sf_point_to_rcoords = l.sf_R;
  dev->ls_to_subfacet_to_inout = l.sf_state;
  dev->state_pitch_bgcell = l.bg_pitch;
  dev->state_pitch_subcell = l.sc_pitch;
  dev->state_pitch_subfacet = l.sf_pitch;
  return GECUT_OK;
}

int gecut_result_copy(const gecut_result* res, const gecut_buffers* host) {
  if (!res || !host) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = res->ctx;
  DeviceGuard guard(ctx->device);
  const GcLayout& l = res->lay;
  const gecut_sizes& s = res->sizes;
  const int D = s.D, L = s.num_levelsets;
  GC_CHECK(copy_rows_to_host(ctx, host->ls_to_bgcell_to_inoutcut, l.bg_state, l.bg_pitch, s.num_bgcells, L));
  GC_CHECK(copy_to_host(ctx, host->cutcell_to_cell, l.cutcells, s.num_cutcells));
  GC_CHECK(copy_to_host(ctx, host->sc_cell_to_points, l.sc_c2p, s.num_subcells * (D + 1)));
  GC_CHECK(copy_to_host(ctx, host->sc_cell_to_bgcell, l.sc_c2bg, s.num_subcells));
  GC_CHECK(copy_to_host(ctx, host->sc_point_to_coords, l.sc_X, s.num_subcell_points * D));
  GC_CHECK(copy_to_host(ctx, host->sc_point_to_rcoords, l.sc_R, s.num_subcell_points * D));
  if (l.sc_state) GC_CHECK(copy_rows_to_host(ctx, host->ls_to_subcell_to_inout, l.sc_state, l.sc_pitch, s.num_subcells, L));
  GC_CHECK(copy_to_host(ctx, host->sf_facet_to_points, l.sf_c2p, s.num_subfacets * D));
  GC_CHECK(copy_to_host(ctx, host->sf_facet_to_bgcell, l.sf_c2bg, s.num_subfacets));
  GC_CHECK(copy_to_host(ctx, host->sf_facet_to_normal, l.sf_N, s.num_subfacets * D));
  GC_CHECK(copy_to_host(ctx, host->sf_point_to_coords, l.sf_X, s.num_subfacet_points * D));
  GC_CHECK(copy_to_host(ctx, host->sf_point_to_rcoords, l.sf_R, s.num_subfacet_points * D));
  if (l.sf_state) GC_CHECK(copy_rows_to_host(ctx, host->ls_to_subfacet_to_inout, l.sf_state, l.sf_pitch, s.num_subfacets, L));
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  return GECUT_OK;
}

static int eval_rows_to_host(const gecut_result* res, const int32_t* program, int len, const int8_t* rows, int64_t pitch,
                             int64_t n, int8_t* host_out) {
  if (!res) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = res->ctx;
  if (!host_out) {
    ctx->err = "NULL output";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  GcProgram p{};
  GC_CHECK(make_program(ctx, program, len, res->L, &p));
  if (n == 0) return GECUT_OK;
  int8_t* d = nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&d, (size_t)n));
  int st = gc_eval_rows(ctx, rows, pitch, n, p, d);
  if (st == GECUT_OK) {
    cudaError_t e = cudaMemcpyAsync(host_out, d, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      ctx->err = cudaGetErrorString(e);
      st = GECUT_ERR_CUDA;
    }
  }
  gc_free(ctx, d);
  return st;
}

int gecut_bgcell_to_inoutcut(const gecut_result* res, const int32_t* program, int len, int8_t* host_out) {
  if (!res) return GECUT_ERR_INVALID;
  return eval_rows_to_host(res, program, len, res->lay.bg_state, res->lay.bg_pitch, res->sizes.num_bgcells, host_out);
}

int gecut_subcell_to_inout(const gecut_result* res, const int32_t* program, int len, int8_t* host_out) {
  if (!res) return GECUT_ERR_INVALID;
  return eval_rows_to_host(res, program, len, res->lay.sc_state, res->lay.sc_pitch, res->sizes.num_subcells, host_out);
}

int gecut_select_cells(const gecut_result* res, int kind, const int32_t* program, int len, int side,
                       gecut_selection** out) {
  if (!res || !out) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = res->ctx;
  *out = nullptr;
  if (kind == GECUT_SEL_BOUNDARY || kind < 0 || kind > GECUT_SEL_BGONLY) {
    ctx->err = "invalid selection kind";
    return GECUT_ERR_INVALID;
  }
  if (side != GECUT_IN && side != GECUT_OUT) {
    ctx->err = "side must be GECUT_IN or GECUT_OUT";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  if (res->free_requested) {
    ctx->err = "the result has been freed";
    return GECUT_ERR_INVALID;
  }
  gecut_selection* sel = new (std::nothrow) gecut_selection();
  if (!sel) return GECUT_ERR_OOM;
  sel->res = res;
  res->live_selections++;
  sel->kind = kind;
  sel->side = side;
  const int64_t launches0 = ctx->launches;
  int st = make_program(ctx, program, len, res->L, &sel->prog);
  if (st == GECUT_OK) st = attach_program_rows(ctx, res, &sel->prog);
  if (st == GECUT_OK) st = gc_select_cells(ctx, sel);
  // selections answered from the tallies of the cut (no kernel queued by this call) do not wait for earlier
  // stream-ordered work of the context (the device-resident Laplacian)
  if (st == GECUT_OK && ctx->launches != launches0 && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = GECUT_ERR_CUDA;
  if (st != GECUT_OK) {
    gecut_selection_free(sel);
    return st;
  }
  *out = sel;
  return GECUT_OK;
}

int gecut_select_boundary(const gecut_result* res, const int32_t* program1, int len1, const int32_t* program2, int len2,
                          gecut_selection** out) {
  if (!res || !out) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = res->ctx;
  *out = nullptr;
  DeviceGuard guard(ctx->device);
  if (res->free_requested) {
    ctx->err = "the result has been freed";
    return GECUT_ERR_INVALID;
  }
  gecut_selection* sel = new (std::nothrow) gecut_selection();
  if (!sel) return GECUT_ERR_OOM;
  sel->res = res;
  res->live_selections++;
  sel->kind = GECUT_SEL_BOUNDARY;
  const int64_t launches0 = ctx->launches;
  int st = make_program(ctx, program1, len1, res->L, &sel->prog);
  if (st == GECUT_OK) st = attach_boundary_lut(ctx, &sel->prog);
  if (st == GECUT_OK && program2) {
    st = make_program(ctx, program2, len2, res->L, &sel->prog2);
    sel->has_prog2 = true;
  }
  if (st == GECUT_OK) st = gc_select_boundary(ctx, sel);
  if (st == GECUT_OK && ctx->launches != launches0 && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = GECUT_ERR_CUDA;
  if (st != GECUT_OK) {
    gecut_selection_free(sel);
    return st;
  }
  *out = sel;
  return GECUT_OK;
}

int gecut_selection_free(gecut_selection* sel) {
  if (!sel) return GECUT_OK;
  gecut_ctx* ctx = sel->res->ctx;
  DeviceGuard guard(ctx->device);
  gc_free(ctx, sel->sub_ids);
  gc_free(ctx, sel->orientation);
  gc_free(ctx, sel->bg_ids);
  gc_free(ctx, sel->K_cut);
  gecut_result* res = const_cast<gecut_result*>(sel->res);
  delete sel;
  if (--res->live_selections == 0 && res->free_requested) return gecut_result_free(res);
  return GECUT_OK;
}

int gecut_selection_sizes(const gecut_selection* sel, int64_t* n_sub, int64_t* n_bg) {
  if (!sel) return GECUT_ERR_INVALID;
  if (n_sub) *n_sub = sel->n_sub;
  if (n_bg) *n_bg = sel->n_bg;
  return GECUT_OK;
}

int gecut_selection_copy(const gecut_selection* sel_c, int32_t* sub_ids, int8_t* orientation, int32_t* bg_ids) {
  if (!sel_c) return GECUT_ERR_INVALID;
  gecut_selection* sel = const_cast<gecut_selection*>(sel_c);
  gecut_ctx* ctx = sel->res->ctx;
  DeviceGuard guard(ctx->device);
  if (sub_ids || orientation) GC_CHECK(gc_materialize_sub_ids(ctx, sel));
  GC_CHECK(copy_to_host(ctx, sub_ids, sel->sub_ids, sel->n_sub));
  if (orientation && sel->kind != GECUT_SEL_BOUNDARY) {
    ctx->err = "orientation is only defined for boundary selections";
    return GECUT_ERR_INVALID;
  }
  GC_CHECK(copy_to_host(ctx, orientation, sel->orientation, sel->n_sub));
  if (bg_ids && sel->n_bg > 0) {
    GC_CHECK(gc_materialize_bg_ids(ctx, sel));
    GC_CHECK(copy_to_host(ctx, bg_ids, sel->bg_ids, sel->n_bg));
  }
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  return GECUT_OK;
}

int gecut_selection_gather(const gecut_selection* sel, int32_t* to_points, int32_t* to_bgcell, double* normals) {
  if (!sel) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = sel->res->ctx;
  DeviceGuard guard(ctx->device);
  const int64_t n = sel->n_sub;
  if (n == 0) return GECUT_OK;
  const int D = sel->res->grid.D;
  const bool boundary = sel->kind == GECUT_SEL_BOUNDARY;
  const int stride = boundary ? D : D + 1;
  if (normals && !boundary) {
    ctx->err = "normals are only defined for boundary selections";
    return GECUT_ERR_INVALID;
  }
  int32_t *d_tp = nullptr, *d_tb = nullptr;
  double* d_n = nullptr;
  int st = GECUT_OK;
  if (to_points) st = gc_malloc(ctx, (void**)&d_tp, sizeof(int32_t) * n * stride);
  if (st == GECUT_OK && to_bgcell) st = gc_malloc(ctx, (void**)&d_tb, sizeof(int32_t) * n);
  if (st == GECUT_OK && normals) st = gc_malloc(ctx, (void**)&d_n, sizeof(double) * n * D);
  if (st == GECUT_OK) st = gc_selection_gather(ctx, sel, d_tp, d_tb, d_n);
  if (st == GECUT_OK) st = copy_to_host(ctx, to_points, d_tp, n * stride);
  if (st == GECUT_OK) st = copy_to_host(ctx, to_bgcell, d_tb, n);
  if (st == GECUT_OK) st = copy_to_host(ctx, normals, d_n, n * D);
  if (st == GECUT_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = GECUT_ERR_CUDA;
  gc_free(ctx, d_tp);
  gc_free(ctx, d_tb);
  gc_free(ctx, d_n);
  return st;
}

int gecut_integrate_measure(const gecut_selection* sel, double* values, double* total) {
  if (!sel || !total) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = sel->res->ctx;
  DeviceGuard guard(ctx->device);
  double* d_vals = nullptr;
  const int64_t n = sel->n_sub;
  if (values && n > 0) GC_CHECK(gc_malloc(ctx, (void**)&d_vals, sizeof(double) * n));
  int st = gc_integrate_measure(ctx, sel, d_vals, total);
  if (st == GECUT_OK && values && n > 0) {
    st = copy_to_host(ctx, values, d_vals, n);
    if (st == GECUT_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = GECUT_ERR_CUDA;
  }
  gc_free(ctx, d_vals);
  return st;
}

int gecut_integrate_laplacian(const gecut_selection* sel_c, double* K_cut, double* K_bg) {
  if (!sel_c) return GECUT_ERR_INVALID;
  gecut_selection* sel = const_cast<gecut_selection*>(sel_c);
  gecut_ctx* ctx = sel->res->ctx;
  if (sel->kind != GECUT_SEL_PHYSICAL && sel->kind != GECUT_SEL_CUTONLY) {
    ctx->err = "laplacian integration needs a PHYSICAL or CUT selection";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  GC_CHECK(gc_integrate_laplacian(ctx, sel, K_bg));
  const int nn = sel->res->grid.nv * sel->res->grid.nv;
  // K_cut == NULL: the matrices stay on the device (gecut_selection_device_laplacian) and the call is stream-ordered --
  // it returns once the kernel is queued, so host work of the caller overlaps it; every later call on this context (and
  // anything the caller queues on the context's stream) sees the finished matrices
  if (!K_cut) return GECUT_OK;
  GC_CHECK(copy_to_host(ctx, K_cut, sel->K_cut, sel->res->sizes.num_cutcells * nn));
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  return GECUT_OK;
}

int gecut_cell_measure(const gecut_selection* sel, double* values, int values_on_device) {
  if (!sel || !values) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = sel->res->ctx;
  if (sel->kind != GECUT_SEL_PHYSICAL && sel->kind != GECUT_SEL_CUTONLY && sel->kind != GECUT_SEL_BGONLY) {
    ctx->err = "cell measures need a PHYSICAL, CUT or bg-cell selection";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  const int64_t nc = sel->res->sizes.num_bgcells;
  if (values_on_device) {
    GC_CHECK(gc_cell_measure(ctx, sel, values));
    GC_CUDA(cudaStreamSynchronize(ctx->stream));
    return GECUT_OK;
  }
  double* d = nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&d, sizeof(double) * std::max<int64_t>(nc, 1)));
  int st = gc_cell_measure(ctx, sel, d);
  if (st == GECUT_OK) st = copy_to_host(ctx, values, d, nc);
  if (st == GECUT_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = GECUT_ERR_CUDA;
  gc_free(ctx, d);
  return st;
}

int gecut_selection_device_laplacian(const gecut_selection* sel, double** K_cut_dev) {
  if (!sel || !K_cut_dev) return GECUT_ERR_INVALID;
  *K_cut_dev = sel->K_cut;
  return GECUT_OK;
}

}  // extern "C"
