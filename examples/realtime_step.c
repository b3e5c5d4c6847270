/*
 * Plain-C caller of the C ABI (include/dce.h): the 1 kHz control-loop step the upstream real-time runner performs
 * (/root/reference/README.md:67-83), without Python.  Weights: 14 fp32 state_dict tensors already on the device
 * (params_dev, in the order dce_weights_pack documents).  Input and outputs live in pinned HOST memory; the fused
 * latency kernel reads / writes them in place, so a step is one launch and one stream synchronise — or, with the
 * resident server (dce_latency_server_start), one store to a doorbell word and a spin on the answer word.
 * `realtime_step --gpu [steps]` times both forms on a B200 (random weights; bench.py runs it for latency_b1.c_caller).
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/realtime_step.c \
 *       -L deep_contact_estimator_b200 -ldce_b200 -L /usr/local/cuda/lib64 -lcudart -o realtime_step
 */
#define _POSIX_C_SOURCE 199309L
#include <stdio.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "dce.h"

typedef struct {
    dce_weights *w;
    void *workspace;             /* device, zero-filled once */
    size_t workspace_bytes;
    float *x_host;               /* pinned: [150][54] z-scored window (utils/data_handler.py:55-56) */
    int32_t *cls_host;           /* pinned: class 0..15 */
    uint8_t *bits_host;          /* pinned: 4 contact bits, MSB first = leg 0 (RF) */
    cudaStream_t stream;
} dce_realtime;

int dce_realtime_open(dce_realtime *rt, int device, const float *const params_dev[DCE_NUM_PARAMS]) {
    int rc;
    memset(rt, 0, sizeof *rt);
    if (cudaSetDevice(device) != cudaSuccess) return DCE_ECUDA;
    if (cudaStreamCreate(&rt->stream) != cudaSuccess) return DCE_ECUDA;
    if ((rc = dce_weights_create(&rt->w, device)) != DCE_OK) return rc;
    if ((rc = dce_weights_pack(rt->w, params_dev, rt->stream)) != DCE_OK) return rc;
    rt->workspace_bytes = dce_workspace_bytes(1, DCE_PREC_BF16X3);
    if (cudaMalloc(&rt->workspace, rt->workspace_bytes) != cudaSuccess) return DCE_ECUDA;
    if (cudaMemsetAsync(rt->workspace, 0, rt->workspace_bytes, rt->stream) != cudaSuccess) return DCE_ECUDA;
    if (cudaHostAlloc((void **)&rt->x_host, DCE_WINDOW * DCE_CHANNELS * sizeof(float), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void **)&rt->cls_host, sizeof(int32_t), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void **)&rt->bits_host, DCE_LEGS, cudaHostAllocDefault) != cudaSuccess)
        return DCE_ECUDA;
    return cudaStreamSynchronize(rt->stream) == cudaSuccess ? DCE_OK : DCE_ECUDA;
}

/* classify the window currently in rt->x_host; on return rt->cls_host / rt->bits_host hold the result */
int dce_realtime_step(dce_realtime *rt) {
    int rc = dce_forward(rt->w, rt->x_host, 1, NULL, rt->cls_host, rt->bits_host, rt->workspace, rt->workspace_bytes,
                         DCE_PREC_BF16X3, rt->stream);
    if (rc != DCE_OK) return rc;
    return cudaStreamSynchronize(rt->stream) == cudaSuccess ? DCE_OK : DCE_ECUDA;
}

void dce_realtime_close(dce_realtime *rt) {
    if (rt->x_host) cudaFreeHost(rt->x_host);
    if (rt->cls_host) cudaFreeHost(rt->cls_host);
    if (rt->bits_host) cudaFreeHost(rt->bits_host);
    if (rt->workspace) cudaFree(rt->workspace);
    dce_weights_destroy(rt->w);
    if (rt->stream) cudaStreamDestroy(rt->stream);
    memset(rt, 0, sizeof *rt);
}

/* ---- the resident form: one cooperative kernel serves a step per doorbell (dce_latency_server_start) ---- */
typedef struct {
    dce_realtime *rt;
    dce_latency_ctrl *ctrl;      /* pinned */
} dce_realtime_server;

int dce_realtime_server_open(dce_realtime_server *sv, dce_realtime *rt, double idle_timeout_s) {
    sv->rt = rt;
    if (cudaHostAlloc((void **)&sv->ctrl, sizeof(dce_latency_ctrl), cudaHostAllocDefault) != cudaSuccess) return DCE_ECUDA;
    memset((void *)sv->ctrl, 0, sizeof(dce_latency_ctrl));
    /* results in the control block itself: class, bits and the step number arrive in one 16-byte store */
    int rc = dce_latency_server_start(rt->w, rt->x_host, 1, NULL, (int32_t *)&sv->ctrl->cls0, (uint8_t *)sv->ctrl->bits0, sv->ctrl,
                                      rt->workspace, rt->workspace_bytes, idle_timeout_s, rt->stream);
    if (rc != DCE_OK) return rc;
    while (!sv->ctrl->alive) { }                      /* the kernel is resident once it has said so */
    return DCE_OK;
}

void dce_realtime_server_close(dce_realtime_server *sv) {
    if (!sv->ctrl) return;
    sv->ctrl->quit = 1;
    cudaStreamSynchronize(sv->rt->stream);
    cudaFreeHost((void *)sv->ctrl);
    sv->ctrl = NULL;
}

/* ---- `realtime_step --gpu [steps]`: both forms on random weights and windows, host wall-clock percentiles ---- */
#include <stdlib.h>
#include <time.h>

static double now_us(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
static int cmp_double(const void *a, const void *b) { return (*(const double *)a > *(const double *)b) - (*(const double *)a < *(const double *)b); }
static unsigned lcg_state = 12345u;
static float lcg_uniform(void) { lcg_state = lcg_state * 1664525u + 1013904223u; return (float)(lcg_state >> 8) / 16777216.0f - 0.5f; }

static int run_gpu(int steps) {
    static const size_t numel[DCE_NUM_PARAMS] = {64 * 54 * 3, 64, 64 * 64 * 3, 64, 128 * 64 * 3, 128, 128 * 128 * 3, 128,
                                                 (size_t)2048 * 4736, 2048, (size_t)512 * 2048, 512, 16 * 512, 16};
    static const float fan_in[DCE_NUM_PARAMS] = {162, 162, 192, 192, 192, 192, 384, 384, 4736, 4736, 2048, 2048, 512, 512};
    const float *params[DCE_NUM_PARAMS];
    dce_realtime rt;
    dce_realtime_server sv;
    double *t = (double *)malloc(sizeof(double) * (size_t)steps);
    int rc, i, same = 1;
    for (i = 0; i < DCE_NUM_PARAMS; ++i) {            /* uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)): the scale of torch's default init */
        float *h = (float *)malloc(numel[i] * sizeof(float)), *d = NULL, bound = 1.0f;
        size_t j;
        while (bound * bound * fan_in[i] > 1.0f) bound *= 0.97f;
        for (j = 0; j < numel[i]; ++j) h[j] = 2.0f * bound * lcg_uniform();
        if (cudaMalloc((void **)&d, numel[i] * sizeof(float)) != cudaSuccess ||
            cudaMemcpy(d, h, numel[i] * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return 2;
        free(h);
        params[i] = d;
    }
    if ((rc = dce_realtime_open(&rt, 0, params)) != DCE_OK) { printf("open: %s\n", dce_strerror(rc)); return 2; }
    int32_t *want = (int32_t *)malloc(sizeof(int32_t) * (size_t)steps);
    for (i = 0; i < steps + 20; ++i) {                /* launch-per-step form */
        int k = i - 20, j;
        lcg_state = 777u + (unsigned)(i % 64);
        for (j = 0; j < DCE_WINDOW * DCE_CHANNELS; ++j) rt.x_host[j] = 3.4f * lcg_uniform();
        double t0 = now_us();
        if ((rc = dce_realtime_step(&rt)) != DCE_OK) { printf("step: %s\n", dce_strerror(rc)); return 2; }
        if (k >= 0) { t[k] = now_us() - t0; want[k] = rt.cls_host[0]; }
    }
    qsort(t, (size_t)steps, sizeof(double), cmp_double);
    printf("C caller, launch per step : host p50 %.1f us  p99 %.1f us  (%d steps)\n", t[steps / 2], t[steps * 99 / 100], steps);
    if ((rc = dce_realtime_server_open(&sv, &rt, 2.0)) != DCE_OK) { printf("server: %s\n", dce_strerror(rc)); return 2; }
    for (i = 0; i < steps + 20; ++i) {                /* resident server: write the window, ring the doorbell, spin */
        int k = i - 20, j;
        lcg_state = 777u + (unsigned)(i % 64);
        for (j = 0; j < DCE_WINDOW * DCE_CHANNELS; ++j) rt.x_host[j] = 3.4f * lcg_uniform();
        double t0 = now_us();
        if (dce_latency_server_step(sv.ctrl) != 0) { printf("server retired\n"); return 2; }
        if (k >= 0) { t[k] = now_us() - t0; same &= (want[k] == sv.ctrl->cls0); }
    }
    qsort(t, (size_t)steps, sizeof(double), cmp_double);
    printf("C caller, resident server : host p50 %.1f us  p99 %.1f us  device %.1f us  classes equal to the launch-per-step form: %s\n",
           t[steps / 2], t[steps * 99 / 100], sv.ctrl->device_ns * 1e-3, same ? "yes" : "NO");
    dce_realtime_server_close(&sv);
    {   /* row server: one new 54-float sensor row per step; ring + z-score + classification on the device */
        dce_latency_row_ctrl *rc_ = NULL;
        float *ring = NULL, row[DCE_CHANNELS];
        if (cudaHostAlloc((void **)&rc_, sizeof *rc_, cudaHostAllocDefault) != cudaSuccess ||
            cudaMalloc((void **)&ring, 2 * DCE_WINDOW * DCE_CHANNELS * sizeof(float)) != cudaSuccess ||
            cudaMemset(ring, 0, 2 * DCE_WINDOW * DCE_CHANNELS * sizeof(float)) != cudaSuccess) return 2;
        if ((rc = dce_latency_row_server_start(rt.w, rc_, ring, rt.workspace, rt.workspace_bytes, 2.0, rt.stream)) != DCE_OK) {
            printf("row server: %s\n", dce_strerror(rc)); return 2;
        }
        while (!rc_->alive) { }
        lcg_state = 4242u;
        for (i = 0; i < steps + DCE_WINDOW; ++i) {
            int k = i - DCE_WINDOW, j;
            for (j = 0; j < DCE_CHANNELS; ++j) row[j] = 3.4f * lcg_uniform() + 0.1f * (float)j;
            double t0 = now_us();
            if (dce_latency_row_server_push(rc_, row, (uint32_t)(i % DCE_WINDOW), (uint32_t)(i + 1)) != 0) { printf("row server retired\n"); return 2; }
            if (k >= 0) t[k] = now_us() - t0;
        }
        qsort(t, (size_t)steps, sizeof(double), cmp_double);
        printf("C caller, row server      : host p50 %.1f us  p99 %.1f us  device %.1f us  (last class %d)\n",
               t[steps / 2], t[steps * 99 / 100], rc_->device_ns * 1e-3, (int)rc_->cls0);
        rc_->chunk[18][0] = 1;                             /* quit */
        cudaStreamSynchronize(rt.stream);
        cudaFreeHost((void *)rc_);
        cudaFree(ring);
    }
    dce_realtime_close(&rt);
    free(t); free(want);
    return same ? 0 : 3;
}

int main(int argc, char **argv) {
    if (argc > 1 && !strcmp(argv[1], "--gpu")) return run_gpu(argc > 2 ? atoi(argv[2]) : 2000);

    /* no GPU needed for this part: the library loads and validates its arguments */
    printf("libdce_b200 version %d; dce_forward(NULL, ...) -> %s\n", dce_version(),
           dce_strerror(dce_forward(NULL, NULL, 1, NULL, NULL, NULL, NULL, 0, DCE_PREC_BF16X3, NULL)));
    printf("workspace for one window: %zu bytes\n", dce_workspace_bytes(1, DCE_PREC_BF16X3));
    return dce_version() == DCE_VERSION ? 0 : 1;
}
