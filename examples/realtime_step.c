/*
 * Plain-C caller of the C ABI (include/dce.h): the 1 kHz control-loop step the upstream real-time runner performs
 * (/root/reference/README.md:67-83), without Python.  Weights: 14 fp32 state_dict tensors already on the device
 * (params_dev, in the order dce_weights_pack documents).  Input and outputs live in pinned HOST memory; the fused
 * latency kernel reads / writes them in place, so a step is one launch and one stream synchronise.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/realtime_step.c \
 *       -L deep_contact_estimator_b200 -ldce_b200 -L /usr/local/cuda/lib64 -lcudart -o realtime_step
 */
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

int main(void) {
    /* no GPU needed for this part: the library loads and validates its arguments */
    printf("libdce_b200 version %d; dce_forward(NULL, ...) -> %s\n", dce_version(),
           dce_strerror(dce_forward(NULL, NULL, 1, NULL, NULL, NULL, NULL, 0, DCE_PREC_BF16X3, NULL)));
    printf("workspace for one window: %zu bytes\n", dce_workspace_bytes(1, DCE_PREC_BF16X3));
    return dce_version() == DCE_VERSION ? 0 : 1;
}
