// gecut_comm.cu -- the multi-GPU part of the C ABI: NCCL collectives enqueued on the context's stream.
//
// Reference pattern replaced (src/Distributed/DistributedDiscretizations.jl):
//   cut(cutter, bgmodel::DistributedDiscreteModel, ...)   :33-40    every rank cuts its own part of the model
//   isconsistent_bgcell_to_inoutcut                        :136-174  `consistent!` ghost update + compare + reduce(&)
//   sum(∫...) over a DistributedTriangulation              (GridapDistributed) all-reduce of the local integrals
//
// z-slab partition: the cut itself needs no data-path collective.  What crosses ranks is
//   * the node plane two neighbouring slabs share, for level sets that exist only as slab-local nodal vectors
//     (ncclSend / ncclRecv of (nx+1)(ny+1) doubles per level set and rank boundary, grouped);
//   * a handful of doubles per step (global integrals and counts): ncclAllReduce(sum);
//   * the consistency flag: ncclAllReduce(min);
//   * per-rank counts for the offsets of the rank-major global layout: ncclAllGather.
// Every collective is enqueued on ctx->stream (stream-ordered with the kernels that produce / consume the data); only the
// calls that return a value to the host synchronise.  A Julia host reaches all of this through ccall alone: rank 0 asks
// for a unique id, distributes the 128 bytes over whatever it already has (MPI.Bcast in GridapDistributed) and every rank
// calls gecut_comm_init.
#include <nccl.h>

#include <cstring>

#include "gecut_internal.cuh"

static_assert(sizeof(ncclUniqueId) == GECUT_COMM_ID_BYTES, "GECUT_COMM_ID_BYTES must be sizeof(ncclUniqueId)");

#define GC_NCCL(call)                                                        \
  do {                                                                       \
    ncclResult_t _r = (call);                                                \
    if (_r != ncclSuccess) {                                                 \
      ctx->err = std::string(#call) + ": " + ncclGetErrorString(_r);         \
      return GECUT_ERR_COMM;                                                 \
    }                                                                        \
  } while (0)

namespace {

struct StoreArgs {
  double v[GECUT_COMM_MAX_VALUES];
};

// host numbers -> device buffer without a host-to-device copy: the values travel as kernel arguments
__global__ void k_store_values(const StoreArgs a, int n, double* __restrict__ out) {
  if ((int)threadIdx.x < n) out[threadIdx.x] = a.v[threadIdx.x];
}

// signs (`v > 0`, what compute_case looks at, LookupTables.jl:40-42) of two copies of the same node plane; NaN never equals
__global__ void __launch_bounds__(256)
k_plane_sign_mismatch(const double* __restrict__ mine, const double* __restrict__ theirs, int64_t n,
                      int* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  if (i < n) {
    const double a = mine[i], b = theirs[i];
    bad = (a > 0.0) != (b > 0.0) || (a != a) || (b != b);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicMin(flag, 0);
}

__global__ void k_set_int(int* p, int v) { *p = v; }

int need_comm(gecut_ctx* ctx) {
  if (!ctx->comm) {
    ctx->err = "no communicator on this context (gecut_comm_init)";
    return GECUT_ERR_INVALID;
  }
  return GECUT_OK;
}

}  // namespace

// ghost layers of the slab-wise aggregation (gecut_aggregate.cu): `send_lo` goes to rank-1 and `recv_lo` comes from it,
// `send_hi` / `recv_hi` likewise with rank+1; one group, stream-ordered
int gc_comm_exchange_layers(gecut_ctx* ctx, const void* send_lo, void* recv_lo, const void* send_hi, void* recv_hi, size_t bytes) {
  GC_CHECK(need_comm(ctx));
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  const int r = ctx->comm_rank, w = ctx->comm_world;
  GC_NCCL(ncclGroupStart());
  if (r > 0) {
    GC_NCCL(ncclSend(send_lo, bytes, ncclUint8, r - 1, comm, ctx->stream));
    GC_NCCL(ncclRecv(recv_lo, bytes, ncclUint8, r - 1, comm, ctx->stream));
  }
  if (r + 1 < w) {
    GC_NCCL(ncclSend(send_hi, bytes, ncclUint8, r + 1, comm, ctx->stream));
    GC_NCCL(ncclRecv(recv_hi, bytes, ncclUint8, r + 1, comm, ctx->stream));
  }
  GC_NCCL(ncclGroupEnd());
  return GECUT_OK;
}

int gc_comm_allreduce_sum_u32(gecut_ctx* ctx, uint32_t* value_dev) {
  GC_CHECK(need_comm(ctx));
  GC_NCCL(ncclAllReduce(value_dev, value_dev, 1, ncclUint32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  return GECUT_OK;
}

extern "C" {

int gecut_comm_unique_id(void* id) {
  if (!id) return GECUT_ERR_INVALID;
  ncclUniqueId u;
  if (ncclGetUniqueId(&u) != ncclSuccess) return GECUT_ERR_COMM;
  memcpy(id, &u, sizeof(u));
  return GECUT_OK;
}

int gecut_comm_init(gecut_ctx* ctx, const void* id, int rank, int world) {
  if (!ctx) return GECUT_ERR_INVALID;
  if (!id || world < 1 || rank < 0 || rank >= world) {
    ctx->err = "invalid communicator arguments";
    return GECUT_ERR_INVALID;
  }
  if (ctx->comm) {
    ctx->err = "the context already has a communicator";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t comm = nullptr;
  GC_NCCL(ncclCommInitRank(&comm, world, u, rank));
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  GC_CUDA(cudaMalloc((void**)&ctx->comm_flag, sizeof(int) * 2));
  return GECUT_OK;
}

int gecut_comm_free(gecut_ctx* ctx) {
  if (!ctx) return GECUT_ERR_INVALID;
  if (!ctx->comm) return GECUT_OK;
  DeviceGuard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ncclCommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  ctx->comm_world = 1;
  ctx->comm_rank = 0;
  cudaFree(ctx->comm_flag);
  ctx->comm_flag = nullptr;
  gc_free(ctx, ctx->comm_plane);
  ctx->comm_plane = nullptr;
  ctx->comm_plane_bytes = 0;
  return GECUT_OK;
}

int gecut_comm_rank(const gecut_ctx* ctx, int* rank, int* world) {
  if (!ctx) return GECUT_ERR_INVALID;
  if (rank) *rank = ctx->comm_rank;
  if (world) *world = ctx->comm_world;
  return GECUT_OK;
}

int gecut_comm_exchange_ghost_planes(gecut_ctx* ctx, double* const* slab_values, int num_vectors, int64_t layer_nodes,
                                     int64_t num_planes) {
  if (!ctx) return GECUT_ERR_INVALID;
  GC_CHECK(need_comm(ctx));
  if (!slab_values || num_vectors < 0 || layer_nodes <= 0 || num_planes < 2) {
    ctx->err = "invalid ghost-plane arguments (a slab holds at least two node planes)";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  const int r = ctx->comm_rank, w = ctx->comm_world;
  if (w == 1 || num_vectors == 0) return GECUT_OK;
  // rank r owns planes k0..k1-1 of its (k1-k0+1)-plane buffer; the last plane is the first plane of rank r+1
  GC_NCCL(ncclGroupStart());
  for (int i = 0; i < num_vectors; ++i) {
    if (!slab_values[i]) continue;
    if (r > 0) GC_NCCL(ncclSend(slab_values[i], (size_t)layer_nodes, ncclDouble, r - 1, comm, ctx->stream));
    if (r < w - 1)
      GC_NCCL(ncclRecv(slab_values[i] + (num_planes - 1) * layer_nodes, (size_t)layer_nodes, ncclDouble, r + 1, comm,
                       ctx->stream));
  }
  GC_NCCL(ncclGroupEnd());
  ctx->launches++;  // one grouped NCCL kernel
  return GECUT_OK;
}

int gecut_comm_check_consistency(gecut_ctx* ctx, const double* const* slab_values, int num_vectors, int64_t layer_nodes,
                                 int64_t num_planes, int* consistent) {
  if (!ctx) return GECUT_ERR_INVALID;
  GC_CHECK(need_comm(ctx));
  if (!consistent || num_vectors < 0 || (num_vectors > 0 && (!slab_values || layer_nodes <= 0 || num_planes < 2))) {
    ctx->err = "invalid consistency-check arguments";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  const int r = ctx->comm_rank, w = ctx->comm_world;
  int* flag = ctx->comm_flag;
  k_set_int<<<1, 1, 0, ctx->stream>>>(flag, 1);
  if (w > 1 && num_vectors > 0) {
    // the owner's copy of the shared plane (first plane of rank r+1) next to the copy this rank cut with (its last plane)
    const size_t need = sizeof(double) * (size_t)layer_nodes * (size_t)num_vectors;
    if (ctx->comm_plane_bytes < need) {
      gc_free(ctx, ctx->comm_plane);
      ctx->comm_plane = nullptr;
      ctx->comm_plane_bytes = 0;
      GC_CHECK(gc_malloc(ctx, (void**)&ctx->comm_plane, need));
      ctx->comm_plane_bytes = need;
    }
    GC_NCCL(ncclGroupStart());
    for (int i = 0; i < num_vectors; ++i) {
      if (!slab_values[i]) continue;
      if (r > 0) GC_NCCL(ncclSend(slab_values[i], (size_t)layer_nodes, ncclDouble, r - 1, comm, ctx->stream));
      if (r < w - 1)
        GC_NCCL(ncclRecv(ctx->comm_plane + (size_t)i * layer_nodes, (size_t)layer_nodes, ncclDouble, r + 1, comm, ctx->stream));
    }
    GC_NCCL(ncclGroupEnd());
    ctx->launches++;
    if (r < w - 1) {
      for (int i = 0; i < num_vectors; ++i) {
        if (!slab_values[i]) continue;
        GC_LAUNCH(ctx, "k_plane_sign_mismatch",
                  k_plane_sign_mismatch<<<(unsigned)gc_div_up(layer_nodes, 256), 256, 0, ctx->stream>>>(
                      slab_values[i] + (num_planes - 1) * layer_nodes, ctx->comm_plane + (size_t)i * layer_nodes,
                      layer_nodes, flag));
      }
    }
  }
  if (w > 1) {
    GC_NCCL(ncclAllReduce(flag, flag, 1, ncclInt, ncclMin, comm, ctx->stream));  // reduce(&, ...)
    ctx->launches++;
  }
  int* h = (int*)ctx->h_pinned;
  GC_CUDA(cudaMemcpyAsync(h, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  *consistent = h[0];
  return GECUT_OK;
}

int gecut_comm_allreduce_sum(gecut_ctx* ctx, const double* host_values, int n, double* result_dev) {
  if (!ctx) return GECUT_ERR_INVALID;
  GC_CHECK(need_comm(ctx));
  if (!result_dev || n < 1 || n > GECUT_COMM_MAX_VALUES) {
    ctx->err = "invalid all-reduce arguments";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  if (host_values) {
    StoreArgs a{};
    for (int i = 0; i < n; ++i) a.v[i] = host_values[i];
    GC_LAUNCH(ctx, "k_store_values", k_store_values<<<1, 32, 0, ctx->stream>>>(a, n, result_dev));
  }
  if (ctx->comm_world > 1) {
    GC_NCCL(ncclAllReduce(result_dev, result_dev, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    ctx->launches++;
  }
  return gc_launch_check(ctx, "gecut_comm_allreduce_sum");
}

int gecut_comm_fetch(gecut_ctx* ctx, const double* values_dev, int n, double* host_out) {
  if (!ctx || !values_dev || !host_out || n < 1) return GECUT_ERR_INVALID;
  DeviceGuard guard(ctx->device);
  GC_CUDA(cudaMemcpyAsync(host_out, values_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  return GECUT_OK;
}

int gecut_comm_allgather_sizes(gecut_ctx* ctx, const gecut_result* res, gecut_sizes* all_sizes) {
  if (!ctx || !res || !all_sizes) return GECUT_ERR_INVALID;
  GC_CHECK(need_comm(ctx));
  DeviceGuard guard(ctx->device);
  const int w = ctx->comm_world;
  const size_t one = sizeof(gecut_sizes);
  char* buf = nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&buf, one * (size_t)w));
  int st = GECUT_OK;
  cudaError_t e = cudaMemcpyAsync(buf + one * ctx->comm_rank, &res->sizes, one, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && w > 1) {
    ncclResult_t r = ncclAllGather(buf + one * ctx->comm_rank, buf, one, ncclChar, (ncclComm_t)ctx->comm, ctx->stream);
    if (r != ncclSuccess) {
      ctx->err = std::string("ncclAllGather: ") + ncclGetErrorString(r);
      st = GECUT_ERR_COMM;
    }
    ctx->launches++;
  }
  if (st == GECUT_OK && e == cudaSuccess) e = cudaMemcpyAsync(all_sizes, buf, one * (size_t)w, cudaMemcpyDeviceToHost, ctx->stream);
  if (st == GECUT_OK && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (st == GECUT_OK && e != cudaSuccess) {
    ctx->err = std::string("gecut_comm_allgather_sizes: ") + cudaGetErrorString(e);
    st = GECUT_ERR_CUDA;
  }
  gc_free(ctx, buf);
  return st;
}

}  // extern "C"
