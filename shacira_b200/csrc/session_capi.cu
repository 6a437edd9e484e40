// session_capi.cu -- host-buffer sessions: the fused latent fwd + bwd step for a caller whose arrays live in HOST
// memory (a non-torch host, or the end-to-end measurement), pipelined two steps deep.
//
// shacira_latent_step_host (capi.cu) is one synchronous call per step: it re-uploads coordinates, table and decoder,
// re-bins the points and serialises upload -> kernels -> download. A session keeps what does not change on the
// device -- the coordinate set and its spatial plan (an image fit has static coordinates, image_trainer.py:234-266) --
// and runs each step on three streams (upload / compute / download) over two slots of device buffers, so that the
// upstream-gradient upload of step i+1 and the feature download of step i use both PCIe directions at once while the
// kernels of either run. Results are complete after shacira_host_session_wait(slot).
#include <cstdlib>

#include "capi_internal.h"

using namespace shacira;

struct shacira_host_session {
    int32_t dim, L, bw, C, F, per_level, nA, device;
    int64_t n, T;
    int32_t first[SHACIRA_MAX_LEVELS], res[SHACIRA_MAX_LEVELS];
    shacira_plan_t* plan;
    char* block;          // one device allocation
    float *coords, *lat, *A, *shift;
    float *gout[2], *feats[2], *glat[2], *gdec[2];   // gdec: [L*C*F] grad_A then [L*F] grad_shift
    cudaStream_t up, comp, down;
    cudaEvent_t ev_up[2], ev_fwd[2], ev_bwd[2], ev_feats_down[2], ev_done[2], ev_table;
    int64_t steps;
    int32_t round_flag, have_shift, have_coords, have_table, issued[2];
};

namespace {
size_t al(size_t b) { return (b + 255) & ~(size_t)255; }
}

extern "C" {

int shacira_host_session_create(int32_t dim, int64_t n, int64_t table_rows, const int32_t* first_idx,
                                const int32_t* resolutions, int32_t num_lods, int32_t codebook_bitwidth,
                                int32_t latent_dim, int32_t feature_dim, int32_t per_level,
                                shacira_host_session_t** out) {
    if (!out) return fail(SHACIRA_ERR_INVALID_ARGUMENT, "session output pointer is NULL");
    *out = nullptr;
    if (n <= 0 || table_rows <= 0) return fail(SHACIRA_ERR_INVALID_ARGUMENT, "n and table_rows must be positive");
    LevelParams lp;
    int rc = build_levels(dim, first_idx, resolutions, num_lods, codebook_bitwidth, lp);
    if (rc) return rc;
    for (int l = 0; l < num_lods; ++l)
        if ((int64_t)lp.first[l] + lp.rows[l] > table_rows)
            return fail(SHACIRA_ERR_INVALID_ARGUMENT, "level %d ends past table_rows", l);
    if (latent_dim != 1 && latent_dim != 2 && latent_dim != 4)
        return fail(SHACIRA_ERR_UNSUPPORTED, "latent_dim %d not in {1,2,4}", latent_dim);
    if (feature_dim != 1 && feature_dim != 2 && feature_dim != 4 && feature_dim != 8)
        return fail(SHACIRA_ERR_UNSUPPORTED, "feature_dim %d not in {1,2,4,8}", feature_dim);
    shacira_host_session* s = new (std::nothrow) shacira_host_session();
    if (!s) return fail(SHACIRA_ERR_CUDA, "out of host memory");
    memset(s, 0, sizeof(*s));
    s->dim = dim; s->n = n; s->T = table_rows; s->L = num_lods; s->bw = codebook_bitwidth;
    s->C = latent_dim; s->F = feature_dim; s->per_level = per_level ? 1 : 0; s->nA = per_level ? num_lods : 1;
    for (int l = 0; l < num_lods; ++l) { s->first[l] = lp.first[l]; s->res[l] = lp.res[l]; }
    cudaGetDevice(&s->device);
    const size_t LF = (size_t)num_lods * feature_dim;
    const size_t b_coords = al(4 * (size_t)n * dim), b_lat = al(4 * (size_t)table_rows * latent_dim);
    const size_t b_A = al(4 * (size_t)s->nA * latent_dim * feature_dim), b_shift = al(4 * (size_t)s->nA * feature_dim);
    const size_t b_rows = al(4 * (size_t)n * LF), b_dec = al(4 * (size_t)num_lods * (latent_dim * feature_dim + feature_dim));
    const size_t total = b_coords + b_lat + b_A + b_shift + 2 * (2 * b_rows + b_lat + b_dec);
    cudaError_t e = cudaMalloc((void**)&s->block, total);
    if (e != cudaSuccess) {
        delete s;
        return fail(SHACIRA_ERR_CUDA, "cudaMalloc(%zu) for the session: %s", total, cudaGetErrorString(e));
    }
    char* p = s->block;
    s->coords = (float*)p; p += b_coords;
    s->lat = (float*)p; p += b_lat;
    s->A = (float*)p; p += b_A;
    s->shift = (float*)p; p += b_shift;
    for (int k = 0; k < 2; ++k) {
        s->gout[k] = (float*)p; p += b_rows;
        s->feats[k] = (float*)p; p += b_rows;
        s->glat[k] = (float*)p; p += b_lat;
        s->gdec[k] = (float*)p; p += b_dec;
    }
    bool ok = cudaStreamCreateWithFlags(&s->up, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&s->comp, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&s->down, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&s->ev_table, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < 2 && ok; ++k)
        ok = cudaEventCreateWithFlags(&s->ev_up[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s->ev_fwd[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s->ev_bwd[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s->ev_feats_down[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s->ev_done[k], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        shacira_host_session_destroy(s);
        return fail(SHACIRA_ERR_CUDA, "session: stream / event creation failed");
    }
    *out = s;
    return SHACIRA_OK;
}

int shacira_host_session_destroy(shacira_host_session_t* s) {
    if (!s) return SHACIRA_OK;
    if (s->up) cudaStreamSynchronize(s->up);
    if (s->comp) cudaStreamSynchronize(s->comp);
    if (s->down) cudaStreamSynchronize(s->down);
    if (s->plan) shacira_plan_destroy(s->plan);
    if (s->block) cudaFree(s->block);
    for (int k = 0; k < 2; ++k) {
        if (s->ev_up[k]) cudaEventDestroy(s->ev_up[k]);
        if (s->ev_fwd[k]) cudaEventDestroy(s->ev_fwd[k]);
        if (s->ev_bwd[k]) cudaEventDestroy(s->ev_bwd[k]);
        if (s->ev_feats_down[k]) cudaEventDestroy(s->ev_feats_down[k]);
        if (s->ev_done[k]) cudaEventDestroy(s->ev_done[k]);
    }
    if (s->ev_table) cudaEventDestroy(s->ev_table);
    if (s->up) cudaStreamDestroy(s->up);
    if (s->comp) cudaStreamDestroy(s->comp);
    if (s->down) cudaStreamDestroy(s->down);
    delete s;
    return SHACIRA_OK;
}

int shacira_host_session_set_coords(shacira_host_session_t* s, const float* coords) {
    if (!s || !coords) return fail(SHACIRA_ERR_INVALID_ARGUMENT, "session / coords is NULL");
    // ordered after every step in flight: the compute stream owns the plan
    CUDA_OK(cudaMemcpyAsync(s->coords, coords, 4 * (size_t)s->n * s->dim, cudaMemcpyHostToDevice, s->comp));
    const bool tiled = (s->dim == 2 && s->L % 4 == 0 && s->n >= 32768) || (s->dim == 3 && s->n >= 393216 && s->C <= 2);   // static coordinates: the plan is built once
    int rc = SHACIRA_OK;
    if (tiled)
        rc = s->plan ? shacira_plan_rebuild(s->plan, s->dim, s->coords, s->n, 0, s->comp)
                     : shacira_plan_create(s->dim, s->coords, s->n, 0, s->comp, &s->plan);
    if (rc) return rc;
    s->have_coords = 1;
    return SHACIRA_OK;
}

int shacira_host_session_set_table(shacira_host_session_t* s, const float* latents, const float* A, const float* shift,
                                   int32_t round_flag) {
    if (!s || !latents || !A) return fail(SHACIRA_ERR_INVALID_ARGUMENT, "session / latents / A is NULL");
    // on the compute stream: ordered after the kernels of the steps in flight that still read the table (1.5 MB at the
    // Kodak shape: ~30 us of PCIe, not worth a second copy of the table)
    CUDA_OK(cudaMemcpyAsync(s->lat, latents, 4 * (size_t)s->T * s->C, cudaMemcpyHostToDevice, s->comp));
    CUDA_OK(cudaMemcpyAsync(s->A, A, 4 * (size_t)s->nA * s->C * s->F, cudaMemcpyHostToDevice, s->comp));
    if (shift) CUDA_OK(cudaMemcpyAsync(s->shift, shift, 4 * (size_t)s->nA * s->F, cudaMemcpyHostToDevice, s->comp));
    s->have_shift = shift ? 1 : 0;
    s->round_flag = round_flag ? 1 : 0;
    s->have_table = 1;
    return SHACIRA_OK;
}

int shacira_host_session_step_async(shacira_host_session_t* s, const float* grad_output, float* feats,
                                    float* grad_latents, float* grad_A, float* grad_shift, int32_t* slot_out) {
    if (!s || !grad_output || !feats || !grad_latents) return fail(SHACIRA_ERR_INVALID_ARGUMENT, "session_step: NULL argument");
    if (!s->have_coords || !s->have_table) return fail(SHACIRA_ERR_INVALID_ARGUMENT, "session_step: set_coords / set_table first");
    const int b = (int)(s->steps & 1);
    const size_t LF = (size_t)s->L * s->F, row_bytes = 4 * (size_t)s->n * LF;
    const bool dec = grad_A != nullptr || grad_shift != nullptr;
    const size_t nA_el = (size_t)s->L * s->C * s->F, nS_el = (size_t)s->L * s->F;
    // upload: the slot's gradient buffer is free once the backward of step - 2 has read it
    if (s->steps >= 2) CUDA_OK(cudaStreamWaitEvent(s->up, s->ev_bwd[b], 0));
    CUDA_OK(cudaMemcpyAsync(s->gout[b], grad_output, row_bytes, cudaMemcpyHostToDevice, s->up));
    CUDA_OK(cudaEventRecord(s->ev_up[b], s->up));
    // compute: forward into the slot's feature buffer (free once step - 2's download has read it)
    if (s->steps >= 2) CUDA_OK(cudaStreamWaitEvent(s->comp, s->ev_feats_down[b], 0));
    int rc;
    const float* shift = s->have_shift ? s->shift : nullptr;
    if (s->plan)
        rc = shacira_latent_forward_planned(s->plan, s->lat, s->first, s->res, s->L, s->bw, s->C, s->F, s->round_flag, s->A,
                                            shift, s->per_level, s->feats[b], s->comp);
    else
        rc = shacira_latent_forward(s->dim, s->coords, s->n, s->lat, s->first, s->res, s->L, s->bw, s->C, s->F,
                                    s->round_flag, s->A, shift, s->per_level, s->feats[b], nullptr, s->comp);
    if (rc) return rc;
    CUDA_OK(cudaEventRecord(s->ev_fwd[b], s->comp));
    // download the features while the backward runs
    CUDA_OK(cudaStreamWaitEvent(s->down, s->ev_fwd[b], 0));
    CUDA_OK(cudaMemcpyAsync(feats, s->feats[b], row_bytes, cudaMemcpyDeviceToHost, s->down));
    CUDA_OK(cudaEventRecord(s->ev_feats_down[b], s->down));
    // backward: needs the upload, and the slot's gradient outputs free (step - 2 fully downloaded)
    CUDA_OK(cudaStreamWaitEvent(s->comp, s->ev_up[b], 0));
    if (s->steps >= 2) CUDA_OK(cudaStreamWaitEvent(s->comp, s->ev_done[b], 0));
    float* gA = dec ? s->gdec[b] : nullptr;
    float* gS = dec ? s->gdec[b] + nA_el : nullptr;
    if (dec) CUDA_OK(cudaMemsetAsync(s->gdec[b], 0, 4 * (nA_el + nS_el), s->comp));
    if (s->plan && s->dim == 2)
        rc = shacira_latent_backward_planned(s->plan, s->gout[b], dec ? s->lat : nullptr, s->first, s->res, s->L, s->bw,
                                             s->C, s->F, s->round_flag, s->A, s->per_level, s->T, 1, s->glat[b], gA, gS,
                                             s->comp);
    else if (s->plan && !dec)
        rc = shacira_latent_backward_planned(s->plan, s->gout[b], nullptr, s->first, s->res, s->L, s->bw, s->C, s->F,
                                             s->round_flag, s->A, s->per_level, s->T, 1, s->glat[b], nullptr, nullptr,
                                             s->comp);
    else if (dec)
        return fail(SHACIRA_ERR_UNSUPPORTED, "session_step: decoder gradients need the 2D tiled path (3D: use the planned_z calls)");
    else
        rc = shacira_latent_backward(s->dim, s->coords, s->n, s->gout[b], nullptr, s->first, s->res, s->L, s->bw, s->C,
                                     s->F, s->A, s->per_level, s->T, 1, s->glat[b], nullptr, nullptr, s->comp);
    if (rc) return rc;
    CUDA_OK(cudaEventRecord(s->ev_bwd[b], s->comp));
    CUDA_OK(cudaStreamWaitEvent(s->down, s->ev_bwd[b], 0));
    CUDA_OK(cudaMemcpyAsync(grad_latents, s->glat[b], 4 * (size_t)s->T * s->C, cudaMemcpyDeviceToHost, s->down));
    if (grad_A) CUDA_OK(cudaMemcpyAsync(grad_A, gA, 4 * nA_el, cudaMemcpyDeviceToHost, s->down));
    if (grad_shift) CUDA_OK(cudaMemcpyAsync(grad_shift, gS, 4 * nS_el, cudaMemcpyDeviceToHost, s->down));
    CUDA_OK(cudaEventRecord(s->ev_done[b], s->down));
    if (slot_out) *slot_out = b;
    s->issued[b] = 1;
    s->steps += 1;
    return SHACIRA_OK;
}

int shacira_host_session_wait(shacira_host_session_t* s, int32_t slot) {
    if (!s || (slot != 0 && slot != 1)) return fail(SHACIRA_ERR_INVALID_ARGUMENT, "session_wait: bad argument");
    if (!s->issued[slot]) return SHACIRA_OK;   // nothing was ever issued on this slot
    CUDA_OK(cudaEventSynchronize(s->ev_done[slot]));
    return SHACIRA_OK;
}

}  // extern "C"
