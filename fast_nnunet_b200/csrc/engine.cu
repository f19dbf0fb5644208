// Network program executor + C ABI of the per-patch forward (include/fnnu.h).
// Replaces `self.network(x)` (predict_from_raw_data.py:543,555) for PlainConvUNet /
// ResidualEncoderUNet programs emitted by fast_nnunet_b200/program.py.
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "ops.cuh"

namespace fnnu {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Buf {
  int d[3];
  int c;
  __half* ptr;
  double* stats;    // [max_batch][c][2]
  ChanMeta* meta;   // [c] (device)
  size_t nvox;
};

struct Op {
  int kind;
  ConvArgs conv;
  EltArgs elt;
  bool umma_ok;
  bool tconv_ok;     // conv_tconv_umma.cu takes this transposed conv
  bool s2_ok;        // conv_s2_umma.cu takes this stride-2 conv
};

}  // namespace fnnu

using namespace fnnu;

struct fnnu_engine {
  std::vector<Buf> bufs;
  std::vector<Op> ops;
  int max_batch;
  int backend;
  double* stats_base;
  size_t stats_bytes;
  int last_total, last_umma;
  int profile_op = -1;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_valid = false;
};

namespace {

struct Plan {
  size_t param_bytes = 0, workspace_bytes = 0;
  std::vector<size_t> buf_off, stat_off, meta_off;
  std::vector<size_t> w_direct_off, w_umma_off, bias_off, staging_op_off;
  std::vector<size_t> w_umma_bytes;
  size_t staging_off = 0, staging_bytes = 0;
  size_t stats_off = 0, stats_bytes = 0;
};

int ntaps_of(const fnnu_op_desc& o) {
  if (o.op == FNNU_OP_TCONV) return o.stride[0] * o.stride[1] * o.stride[2];
  return o.kernel[0] * o.kernel[1] * o.kernel[2];
}

int validate(const fnnu_buffer_desc* bufs, int n_bufs, const fnnu_op_desc* ops, int n_ops, int max_batch) {
  FNNU_CHECK_ARG(bufs && ops && n_bufs > 0 && n_ops > 0, "engine: empty program");
  FNNU_CHECK_ARG(max_batch >= 1 && max_batch <= 65535, "engine: max_batch=%d", max_batch);
  for (int i = 0; i < n_bufs; ++i)
    FNNU_CHECK_ARG(bufs[i].dims[0] > 0 && bufs[i].dims[1] > 0 && bufs[i].dims[2] > 0 && bufs[i].channels > 0,
                   "engine: buffer %d has a non-positive extent", i);
  for (int i = 0; i < n_ops; ++i) {
    const fnnu_op_desc& o = ops[i];
    FNNU_CHECK_ARG(o.op >= FNNU_OP_CONV && o.op <= FNNU_OP_AVGPOOL, "engine: op %d has unknown kind %d", i, o.op);
    FNNU_CHECK_ARG(o.src >= 0 && o.src < n_bufs && o.dst >= 0 && o.dst < n_bufs, "engine: op %d buffer index", i);
    const fnnu_buffer_desc& sb = bufs[o.src];
    const fnnu_buffer_desc& db = bufs[o.dst];
    FNNU_CHECK_ARG(o.cin > 0 && o.cout > 0, "engine: op %d channels", i);
    FNNU_CHECK_ARG(o.src_coff >= 0 && o.src_coff + o.cin <= sb.channels, "engine: op %d source channel range", i);
    FNNU_CHECK_ARG(o.dst_coff >= 0 && o.dst_coff + o.cout <= db.channels, "engine: op %d destination channel range", i);
    for (int a = 0; a < 3; ++a) {
      if (o.op == FNNU_OP_CONV) {
        FNNU_CHECK_ARG(o.kernel[a] == 1 || o.kernel[a] == 3, "engine: op %d kernel %d on axis %d (1 or 3)", i, o.kernel[a], a);
        FNNU_CHECK_ARG(o.stride[a] >= 1 && o.stride[a] <= 2, "engine: op %d stride %d on axis %d", i, o.stride[a], a);
        int pad = (o.kernel[a] - 1) / 2;
        int out = (sb.dims[a] + 2 * pad - o.kernel[a]) / o.stride[a] + 1;
        FNNU_CHECK_ARG(out == db.dims[a], "engine: op %d output extent %d != buffer extent %d on axis %d", i, out, db.dims[a], a);
      } else if (o.op == FNNU_OP_TCONV) {
        FNNU_CHECK_ARG(o.stride[a] >= 1 && o.stride[a] <= 2 && o.kernel[a] == o.stride[a],
                       "engine: op %d transposed conv needs kernel == stride in {1,2}", i);
        FNNU_CHECK_ARG(sb.dims[a] * o.stride[a] == db.dims[a], "engine: op %d transposed conv extent on axis %d", i, a);
      } else if (o.op == FNNU_OP_AVGPOOL) {
        FNNU_CHECK_ARG(o.stride[a] >= 1 && db.dims[a] * o.stride[a] == sb.dims[a], "engine: op %d pooling extent on axis %d", i, a);
      } else {
        FNNU_CHECK_ARG(sb.dims[a] == db.dims[a], "engine: op %d add extents", i);
      }
    }
    if (o.op == FNNU_OP_CONV || o.op == FNNU_OP_TCONV) {
      FNNU_CHECK_ARG(o.weight != nullptr, "engine: op %d has no weight", i);
      FNNU_CHECK_ARG(!o.has_bias || o.bias, "engine: op %d has_bias without bias", i);
      FNNU_CHECK_ARG(!o.has_norm || (o.gamma && o.beta), "engine: op %d has_norm without gamma/beta", i);
      FNNU_CHECK_ARG(!(o.op == FNNU_OP_TCONV && o.has_norm), "engine: op %d norm after transposed conv is not supported", i);
    } else {
      FNNU_CHECK_ARG(o.cin == o.cout, "engine: op %d elementwise op needs cin == cout", i);
    }
    if (o.op == FNNU_OP_ADD_ACT) {
      FNNU_CHECK_ARG(o.src2 >= 0 && o.src2 < n_bufs, "engine: op %d second source", i);
      const fnnu_buffer_desc& s2 = bufs[o.src2];
      FNNU_CHECK_ARG(o.src2_coff >= 0 && o.src2_coff + o.cin <= s2.channels, "engine: op %d second source channel range", i);
      for (int a = 0; a < 3; ++a) FNNU_CHECK_ARG(s2.dims[a] == sb.dims[a], "engine: op %d add extents", i);
    }
  }
  return FNNU_OK;
}

void make_plan(const fnnu_buffer_desc* bufs, int n_bufs, const fnnu_op_desc* ops, int n_ops, int max_batch, Plan& p) {
  size_t w = 0;
  p.buf_off.resize(n_bufs);
  p.stat_off.resize(n_bufs);
  p.meta_off.resize(n_bufs);
  for (int i = 0; i < n_bufs; ++i) {
    p.buf_off[i] = w;
    w += align_up((size_t)max_batch * bufs[i].dims[0] * bufs[i].dims[1] * bufs[i].dims[2] * bufs[i].channels * sizeof(__half), 1024);
  }
  p.stats_off = w;
  for (int i = 0; i < n_bufs; ++i) {
    p.stat_off[i] = w;
    w += align_up((size_t)max_batch * bufs[i].channels * 2 * sizeof(double), 256);
  }
  p.stats_bytes = w - p.stats_off;
  p.workspace_bytes = w;

  size_t q = 0;
  for (int i = 0; i < n_bufs; ++i) {
    p.meta_off[i] = q;
    q += align_up((size_t)bufs[i].channels * sizeof(ChanMeta), 256);
  }
  p.w_direct_off.assign(n_ops, 0);
  p.w_umma_off.assign(n_ops, 0);
  p.w_umma_bytes.assign(n_ops, 0);
  p.bias_off.assign(n_ops, 0);
  p.staging_op_off.assign(n_ops, 0);
  size_t stage = 0;
  for (int i = 0; i < n_ops; ++i) {
    const fnnu_op_desc& o = ops[i];
    if (o.op != FNNU_OP_CONV && o.op != FNNU_OP_TCONV) continue;
    int nt = ntaps_of(o);
    int cout_pad = (o.cout + 15) / 16 * 16;
    p.w_direct_off[i] = q;
    q += align_up((size_t)nt * o.cin * cout_pad * sizeof(float), 256);
    size_t ub = umma_packed_weight_bytes(o.cin, o.cout, nt, o.op == FNNU_OP_TCONV);
    p.w_umma_bytes[i] = ub;
    if (ub) {
      p.w_umma_off[i] = q;
      q += align_up(ub, 1024);
    }
    p.bias_off[i] = q;
    q += align_up((size_t)o.cout * sizeof(float), 256);
    // every op has its own slice of the staging area: the raw fp32 weights of all ops are uploaded back to back and
    // packed on the stream without a host synchronisation per op (a 5-fold ensemble creates 5 engines per model)
    p.staging_op_off[i] = stage;
    stage += align_up((size_t)nt * o.cin * o.cout * sizeof(float), 256);
  }
  p.staging_off = q;
  p.staging_bytes = align_up(stage, 256);
  q += p.staging_bytes;
  p.param_bytes = q;
}

}  // namespace

extern "C" int fnnu_abi_version(void) { return FNNU_ABI_VERSION; }
extern "C" const char* fnnu_last_error(void) { return g_err; }

extern "C" int fnnu_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device");
    return 0;
  }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    set_error("device compute capability %d.x, this library is built for sm_100a only", major);
    return 0;
  }
  return 1;
}

extern "C" int fnnu_engine_sizes(const fnnu_buffer_desc* bufs, int n_bufs, const fnnu_op_desc* ops, int n_ops,
                                 int max_batch, size_t* param_bytes, size_t* workspace_bytes) {
  int rc = validate(bufs, n_bufs, ops, n_ops, max_batch);
  if (rc) return rc;
  Plan p;
  make_plan(bufs, n_bufs, ops, n_ops, max_batch, p);
  if (param_bytes) *param_bytes = p.param_bytes;
  if (workspace_bytes) *workspace_bytes = p.workspace_bytes;
  return FNNU_OK;
}

extern "C" int fnnu_engine_create(const fnnu_buffer_desc* bufs, int n_bufs, const fnnu_op_desc* ops, int n_ops,
                                  int max_batch, void* param_arena, size_t param_bytes, void* workspace,
                                  size_t workspace_bytes, void* stream, fnnu_engine** out) {
  FNNU_CHECK_ARG(out && param_arena && workspace, "engine_create: null pointer");
  int rc = validate(bufs, n_bufs, ops, n_ops, max_batch);
  if (rc) return rc;
  Plan p;
  make_plan(bufs, n_bufs, ops, n_ops, max_batch, p);
  FNNU_CHECK_ARG(param_bytes >= p.param_bytes, "engine_create: param arena %zu < %zu", param_bytes, p.param_bytes);
  FNNU_CHECK_ARG(workspace_bytes >= p.workspace_bytes, "engine_create: workspace %zu < %zu", workspace_bytes, p.workspace_bytes);
  FNNU_CHECK_ARG(((uintptr_t)param_arena % 256 == 0) && ((uintptr_t)workspace % 256 == 0), "engine_create: arenas must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  char* pa = (char*)param_arena;
  char* ws = (char*)workspace;

  fnnu_engine* e = new fnnu_engine();
  e->max_batch = max_batch;
  e->backend = 0;
  e->last_total = e->last_umma = 0;
  e->stats_base = (double*)(ws + p.stats_off);
  e->stats_bytes = p.stats_bytes;
  e->bufs.resize(n_bufs);
  std::vector<std::vector<ChanMeta>> metas(n_bufs);
  for (int i = 0; i < n_bufs; ++i) {
    Buf& b = e->bufs[i];
    for (int a = 0; a < 3; ++a) b.d[a] = bufs[i].dims[a];
    b.c = bufs[i].channels;
    b.nvox = (size_t)b.d[0] * b.d[1] * b.d[2];
    b.ptr = (__half*)(ws + p.buf_off[i]);
    b.stats = (double*)(ws + p.stat_off[i]);
    b.meta = (ChanMeta*)(pa + p.meta_off[i]);
    metas[i].assign(b.c, ChanMeta{1.f, 0.f, 1.f, -1.f});
  }
  // pending transforms left behind by normalised convs
  for (int i = 0; i < n_ops; ++i) {
    const fnnu_op_desc& o = ops[i];
    if (o.op == FNNU_OP_CONV && o.has_norm)
      for (int c = 0; c < o.cout; ++c)
        metas[o.dst][o.dst_coff + c] = ChanMeta{o.gamma[c], o.beta[c], o.act_slope, o.norm_eps};
  }
  for (int i = 0; i < n_bufs; ++i) {
    // pageable source: cudaMemcpyAsync returns once the bytes sit in the driver's staging buffer
    cudaError_t ce = cudaMemcpyAsync(e->bufs[i].meta, metas[i].data(), metas[i].size() * sizeof(ChanMeta), cudaMemcpyHostToDevice, s);
    if (ce != cudaSuccess) {
      set_error("engine_create: meta upload failed: %s", cudaGetErrorString(ce));
      delete e;
      return FNNU_E_CUDA;
    }
  }

  e->ops.resize(n_ops);
  for (int i = 0; i < n_ops; ++i) {
    const fnnu_op_desc& o = ops[i];
    Op& op = e->ops[i];
    op.kind = o.op;
    op.umma_ok = false;
    op.tconv_ok = false;
    op.s2_ok = false;
    const Buf& sb = e->bufs[o.src];
    const Buf& db = e->bufs[o.dst];
    if (o.op == FNNU_OP_CONV || o.op == FNNU_OP_TCONV) {
      ConvArgs& a = op.conv;
      memset(&a, 0, sizeof(a));
      a.src = sb.ptr + o.src_coff;
      a.src_cs = sb.c;
      a.src_stats = sb.stats + (size_t)o.src_coff * 2;
      a.src_meta = sb.meta + o.src_coff;
      a.src_stat_stride = sb.c;
      a.src_inv_count = 1.0 / (double)sb.nvox;
      a.dst = db.ptr + o.dst_coff;
      a.dst_cs = db.c;
      a.dst_stats = o.has_norm ? db.stats + (size_t)o.dst_coff * 2 : nullptr;
      a.dst_stat_stride = db.c;
      a.cin = o.cin;
      a.cout = o.cout;
      a.cout_pad = (o.cout + 15) / 16 * 16;
      a.transposed = o.op == FNNU_OP_TCONV;
      a.ntaps = ntaps_of(o);
      for (int ax = 0; ax < 3; ++ax) {
        a.in_d[ax] = sb.d[ax];
        a.out_d[ax] = db.d[ax];
        a.k[ax] = o.kernel[ax];
        a.s[ax] = o.stride[ax];
        a.pad[ax] = a.transposed ? 0 : (o.kernel[ax] - 1) / 2;
      }
      // upload + pack weights
      float* staging = (float*)(pa + p.staging_off + p.staging_op_off[i]);
      size_t raw = (size_t)a.ntaps * o.cin * o.cout * sizeof(float);
      cudaError_t ce = cudaMemcpyAsync(staging, o.weight, raw, cudaMemcpyHostToDevice, s);
      if (ce != cudaSuccess) {
        set_error("engine_create: weight upload of op %d failed: %s", i, cudaGetErrorString(ce));
        delete e;
        return FNNU_E_CUDA;
      }
      float* wd = (float*)(pa + p.w_direct_off[i]);
      rc = launch_pack_weights_direct(staging, wd, o.cin, o.cout, a.cout_pad, a.ntaps, a.transposed, s);
      if (rc) { delete e; return rc; }
      a.w = wd;
      if (p.w_umma_bytes[i]) {
        void* wu = pa + p.w_umma_off[i];
        a.use_rows = zrows_supported(a) ? 2 : (rows_supported(a) ? 1 : 0);
        rc = a.use_rows == 2 ? launch_pack_weights_zrows(staging, wu, a, s)
             : a.use_rows   ? launch_pack_weights_rows(staging, wu, a, s)
                            : launch_pack_weights_umma(staging, wu, a, s);
        if (rc) { delete e; return rc; }
        a.w_umma = wu;
      }
      float* bd = (float*)(pa + p.bias_off[i]);
      if (o.has_bias) {
        ce = cudaMemcpyAsync(bd, o.bias, (size_t)o.cout * sizeof(float), cudaMemcpyHostToDevice, s);
        a.bias = bd;
      } else {
        a.bias = nullptr;
      }
      if (ce != cudaSuccess) {
        set_error("engine_create: op %d parameter upload failed: %s", i, cudaGetErrorString(ce));
        delete e;
        return FNNU_E_CUDA;
      }
      a.batch = 1;
      op.umma_ok = a.w_umma != nullptr && (a.use_rows || umma_supported(a));
      op.tconv_ok = a.transposed && tconv_umma_supported(a);
      op.s2_ok = s2_umma_supported(a);
    } else {
      EltArgs& a = op.elt;
      memset(&a, 0, sizeof(a));
      a.src = sb.ptr + o.src_coff;
      a.src_cs = sb.c;
      a.src_stats = sb.stats + (size_t)o.src_coff * 2;
      a.src_meta = sb.meta + o.src_coff;
      a.src_stat_stride = sb.c;
      if (o.op == FNNU_OP_ADD_ACT) {
        const Buf& s2 = e->bufs[o.src2];
        a.src2 = s2.ptr + o.src2_coff;
        a.src2_cs = s2.c;
        a.src2_stats = s2.stats + (size_t)o.src2_coff * 2;
        a.src2_meta = s2.meta + o.src2_coff;
        a.src2_stat_stride = s2.c;
      }
      a.dst = db.ptr + o.dst_coff;
      a.dst_cs = db.c;
      for (int ax = 0; ax < 3; ++ax) {
        a.d[ax] = db.d[ax];
        a.s[ax] = o.op == FNNU_OP_AVGPOOL ? o.stride[ax] : 1;
      }
      a.c = o.cin;
      a.inv_count = 1.0 / (double)sb.nvox;
      a.slope = o.act_slope;
      a.batch = 1;
    }
  }
  {
    cudaError_t ce = cudaStreamSynchronize(s);   // ONE synchronisation: the caller's host arrays may go away now
    if (ce != cudaSuccess) {
      set_error("engine_create: parameter upload failed: %s", cudaGetErrorString(ce));
      delete e;
      return FNNU_E_CUDA;
    }
  }
  *out = e;
  return FNNU_OK;
}

extern "C" void fnnu_engine_destroy(fnnu_engine* e) {
  if (!e) return;
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  delete e;
}

extern "C" int fnnu_engine_profile_op(fnnu_engine* e, int op_index) {
  FNNU_CHECK_ARG(e && op_index < (int)e->ops.size(), "profile_op: index %d", op_index);
  if (op_index >= 0 && !e->ev0) {
    FNNU_CUDA(cudaEventCreate(&e->ev0));
    FNNU_CUDA(cudaEventCreate(&e->ev1));
  }
  e->profile_op = op_index;
  e->ev_valid = false;
  return FNNU_OK;
}

extern "C" int fnnu_engine_profile_ms(fnnu_engine* e, float* ms) {
  FNNU_CHECK_ARG(e && ms, "profile_ms: null pointer");
  FNNU_CHECK_ARG(e->ev_valid, "profile_ms: no profiled forward has run");
  FNNU_CUDA(cudaEventSynchronize(e->ev1));
  FNNU_CUDA(cudaEventElapsedTime(ms, e->ev0, e->ev1));
  return FNNU_OK;
}

extern "C" void* fnnu_engine_buffer(fnnu_engine* e, int index) {
  if (!e || index < 0 || index >= (int)e->bufs.size()) return nullptr;
  return e->bufs[index].ptr;
}

extern "C" double* fnnu_engine_stats(fnnu_engine* e, int index) {
  if (!e || index < 0 || index >= (int)e->bufs.size()) return nullptr;
  return e->bufs[index].stats;
}

extern "C" int fnnu_engine_set_backend(fnnu_engine* e, int backend) {
  FNNU_CHECK_ARG(e && (backend == 0 || backend == 1), "set_backend: backend=%d", backend);
  e->backend = backend;
  return FNNU_OK;
}

extern "C" int fnnu_engine_launch_counts(fnnu_engine* e, int* total, int* umma) {
  FNNU_CHECK_ARG(e, "launch_counts: null engine");
  if (total) *total = e->last_total;
  if (umma) *umma = e->last_umma;
  return FNNU_OK;
}

extern "C" int fnnu_engine_forward(fnnu_engine* e, int batch, void* stream) {
  FNNU_CHECK_ARG(e, "forward: null engine");
  FNNU_CHECK_ARG(batch >= 1 && batch <= e->max_batch, "forward: batch %d outside [1, %d]", batch, e->max_batch);
  cudaStream_t s = (cudaStream_t)stream;
  FNNU_CUDA(cudaMemsetAsync(e->stats_base, 0, e->stats_bytes, s));
  int total = 0, umma = 0;
  for (size_t i = 0; i < e->ops.size(); ++i) {
    Op& op = e->ops[i];
    int rc = FNNU_OK;
    const bool prof = (int)i == e->profile_op;
    if (prof) FNNU_CUDA(cudaEventRecord(e->ev0, s));
    if (op.kind == FNNU_OP_CONV || op.kind == FNNU_OP_TCONV) {
      op.conv.batch = batch;
      if (e->backend == 0 && op.tconv_ok) {
        rc = launch_tconv_umma(op.conv, s);
        ++umma;
      } else if (e->backend == 0 && op.s2_ok) {
        rc = launch_conv_s2_umma(op.conv, s);
        ++umma;
      } else if (e->backend == 0 && op.umma_ok && !prefer_cuda_cores(op.conv)) {
        rc = op.conv.use_rows == 2 ? launch_conv_zrows(op.conv, s)
             : op.conv.use_rows   ? launch_conv_rows(op.conv, s)
                                  : launch_conv_umma(op.conv, s);
        ++umma;
      } else if (e->backend == 0 && direct_specialised(op.conv)) {
        rc = launch_conv_specialised(op.conv, s);
      } else {
        rc = launch_conv_direct(op.conv, s);
      }
    } else if (op.kind == FNNU_OP_ADD_ACT) {
      op.elt.batch = batch;
      rc = launch_add_act(op.elt, s);
    } else {
      op.elt.batch = batch;
      rc = launch_avgpool(op.elt, s);
    }
    if (rc) return rc;
    if (prof) {
      FNNU_CUDA(cudaEventRecord(e->ev1, s));
      e->ev_valid = true;
    }
    ++total;
  }
  e->last_total = total;
  e->last_umma = umma;
  return FNNU_OK;
}
