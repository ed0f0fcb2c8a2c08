// dsc_api.cu -- context lifetime and small entry points of the C ABI (include/dsc.h).
#include "dsc_internal.h"
#include <stdlib.h>

void dsc_k1_release(dsc_ctx *ctx);
void dsc_k1f64_release(dsc_ctx *ctx);

extern "C" const char *dsc_version(void) { return DSC_VERSION_STRING; }

extern "C" const char *dsc_last_error(const dsc_ctx *ctx) { return ctx ? ctx->err.c_str() : "ctx is NULL"; }

extern "C" int dsc_create(int device, dsc_ctx **out)
{
    if (!out) return DSC_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return DSC_ERR_CUDA;  // no CPU fallback
    if (device < 0 || device >= count) return DSC_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return DSC_ERR_CUDA;
    dsc_ctx *ctx = new (std::nothrow) dsc_ctx();
    if (!ctx) return DSC_ERR_NOMEM;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return DSC_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_bytes = prop.l2CacheSize;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    bool ok = cudaEventCreate(&ctx->ev0) == cudaSuccess && cudaEventCreate(&ctx->ev1) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_call, cudaEventDisableTiming) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_k, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < DSC_NBUF && ok; ++i)
        ok = cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { dsc_destroy(ctx); return DSC_ERR_CUDA; }
    *out = ctx;
    return DSC_OK;
}

extern "C" int dsc_destroy(dsc_ctx *ctx)
{
    if (!ctx) return DSC_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    dsc_k1_release(ctx);
    dsc_k1f64_release(ctx);
    if (ctx->d_rc) cudaFree(ctx->d_rc);
    if (ctx->d_cos) cudaFree(ctx->d_cos);
    if (ctx->d_sin) cudaFree(ctx->d_sin);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->work_counter) cudaFree(ctx->work_counter);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_call) cudaEventDestroy(ctx->ev_call);
    for (int i = 0; i < DSC_NBUF; ++i) {
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]);
        if (ctx->ev_d2h[i]) cudaEventDestroy(ctx->ev_d2h[i]);
    }
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_k) cudaStreamDestroy(ctx->s_k);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    delete ctx;
    return DSC_OK;
}

extern "C" int dsc_set_timing(dsc_ctx *ctx, int enabled)
{
    if (!ctx) return DSC_ERR_INVALID;
    ctx->timing = enabled != 0;
    ctx->last_ms = -1.0f;
    return DSC_OK;
}
extern "C" int dsc_set_option(dsc_ctx *ctx, const char *name, int64_t value)
{
    if (!ctx || !name) return DSC_ERR_INVALID;
    if (!strcmp(name, "k1_impl") && (value == 0 || value == 1 || value == 3)) { ctx->k1_impl = (int)value; return DSC_OK; }
    if (!strcmp(name, "sc_smem_rows") && value >= -1 && value <= 65536) { ctx->sc_smem_rows = (int)value; return DSC_OK; }
    if (!strcmp(name, "no_defgrid") && value >= 0 && value <= 1) { ctx->no_defgrid = (int)value; return DSC_OK; }
    if (!strcmp(name, "sc_split") && value >= -1 && value <= 1) { ctx->sc_split = (int)value; return DSC_OK; }
    if (!strcmp(name, "sc_item_points") && value >= 1024 && value <= (1 << 30)) { ctx->sc_item_points = (int)value; return DSC_OK; }
    if (!strcmp(name, "sc_window") && value >= -1 && value <= 65536) { ctx->sc_window = (int)value; return DSC_OK; }
    if (!strcmp(name, "sc_cfg") && value >= -1 && value <= 7) { ctx->sc_cfg = (int)value; return DSC_OK; }
    DSC_FAIL(ctx, DSC_ERR_INVALID, std::string("unknown option or bad value: ") + name);
}
extern "C" float dsc_last_kernel_ms(const dsc_ctx *ctx) { return ctx ? ctx->last_ms : -1.0f; }
extern "C" int64_t dsc_kernel_launches(const dsc_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int dsc_host_alloc(dsc_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return DSC_ERR_INVALID;
    *out = nullptr;
    DSC_CUDA(ctx, cudaSetDevice(ctx->device));
    DSC_CUDA(ctx, cudaMallocHost(out, bytes ? bytes : 1));
    return DSC_OK;
}
extern "C" int dsc_host_free(dsc_ctx *ctx, void *p)
{
    if (!ctx) return DSC_ERR_INVALID;
    if (p) DSC_CUDA(ctx, cudaFreeHost(p));
    return DSC_OK;
}

int dsc_enter_stream(dsc_ctx *ctx, cudaStream_t st)
{
    if (ctx->ev_call_set) DSC_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_call, 0));
    return DSC_OK;
}
int dsc_leave_stream(dsc_ctx *ctx, cudaStream_t st)
{
    DSC_CUDA(ctx, cudaEventRecord(ctx->ev_call, st));
    ctx->ev_call_set = 1;
    return DSC_OK;
}

int dsc_reserve_scratch(dsc_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->scratch_cap) return DSC_OK;
    if (ctx->scratch) {
        DSC_CUDA(ctx, cudaDeviceSynchronize());
        cudaFree(ctx->scratch);
        ctx->scratch = nullptr;
        ctx->scratch_cap = 0;
    }
    const size_t cap = bytes + bytes / 4 + 4096;
    DSC_CUDA(ctx, cudaMalloc(&ctx->scratch, cap));
    ctx->scratch_cap = cap;
    return DSC_OK;
}

extern "C" int dsc_set_polar_tables(dsc_ctx *ctx, int num_ring, int num_sector, const double *r_centers,
                                    const double *cos_tab, const double *sin_tab)
{
    if (!ctx) return DSC_ERR_INVALID;
    if (num_ring <= 0 || num_sector <= 0 || (int64_t)num_ring * num_sector > ((int64_t)1 << 26))
        DSC_FAIL(ctx, DSC_ERR_INVALID, "bad table sizes");
    if (!r_centers || !cos_tab || !sin_tab) DSC_FAIL(ctx, DSC_ERR_INVALID, "NULL table");
    DSC_CUDA(ctx, cudaSetDevice(ctx->device));
    DSC_CUDA(ctx, cudaDeviceSynchronize());
    if (ctx->d_rc) cudaFree(ctx->d_rc);
    if (ctx->d_cos) cudaFree(ctx->d_cos);
    if (ctx->d_sin) cudaFree(ctx->d_sin);
    ctx->d_rc = ctx->d_cos = ctx->d_sin = nullptr;
    DSC_CUDA(ctx, cudaMalloc(&ctx->d_rc, sizeof(double) * num_ring));
    DSC_CUDA(ctx, cudaMalloc(&ctx->d_cos, sizeof(double) * num_sector));
    DSC_CUDA(ctx, cudaMalloc(&ctx->d_sin, sizeof(double) * num_sector));
    DSC_CUDA(ctx, cudaMemcpy(ctx->d_rc, r_centers, sizeof(double) * num_ring, cudaMemcpyHostToDevice));
    DSC_CUDA(ctx, cudaMemcpy(ctx->d_cos, cos_tab, sizeof(double) * num_sector, cudaMemcpyHostToDevice));
    DSC_CUDA(ctx, cudaMemcpy(ctx->d_sin, sin_tab, sizeof(double) * num_sector, cudaMemcpyHostToDevice));
    ctx->tab_R = num_ring;
    ctx->tab_S = num_sector;
    return DSC_OK;
}
