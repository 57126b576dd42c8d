// C-ABI, part 3: whole-bag streaming (BASELINE config "offline whole-bag reprocessing").
// A bag is handed over as a list of host batches; batch k+1's clouds cross PCIe on a second stream
// while batch k's kernels run (the resident scan slots are split into two halves that alternate), so
// the end-to-end rate is max(copy, compute) instead of their sum.  One synchronisation at the very end;
// no collective anywhere (ranks own disjoint frame ranges, bag.py gathers the result records once).
#include "vlo_internal.cuh"
#include <cstring>
#include <algorithm>

namespace {

// Kept in the handle between calls: a pinned allocation and a stream creation cost more than a millisecond, a streaming call
// over ten 128-scan batches lasts 35.
struct BagCtx {
    cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
    cudaEvent_t copied[2] = { nullptr, nullptr }, raw_free[2] = { nullptr, nullptr };
    char *pinned = nullptr; size_t pinned_bytes = 0;
    char *d_par = nullptr;          // per half: slot indices (2 arrays) + seeds, uploaded on the COPY stream
    ~BagCtx()
    {
        if (copy_stream) cudaStreamSynchronize(copy_stream);
        if (compute_stream) cudaStreamSynchronize(compute_stream);
        for (int i = 0; i < 2; i++) { if (copied[i]) cudaEventDestroy(copied[i]); if (raw_free[i]) cudaEventDestroy(raw_free[i]); }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (pinned) cudaFreeHost(pinned);
        if (d_par) cudaFree(d_par);
    }
};
// error paths leave work in flight: drain both streams before the caller touches the staging again
struct BagDrain {
    BagCtx &c;
    ~BagDrain() { if (c.copy_stream) cudaStreamSynchronize(c.copy_stream); if (c.compute_stream) cudaStreamSynchronize(c.compute_stream); }
};

// mode 0: scan-to-map of every scan; mode 1: scan-to-scan of the consecutive pairs inside each batch
int bag_run(vlo_handle *h, const vlo_bag_batch *batches, int n_batches, int stride, vlo_result *out, int mode)
{
    if (!h || !batches || !out || n_batches < 1 || stride < 3) return VLO_ERR_INVALID_ARG;
    const int HB = h->cfg.max_scans / 2;
    if (HB < (mode == 1 ? 2 : 1)) { h->err = "whole-bag streaming needs max_scans >= 2 (>= 4 for pairs): the resident slots are double-buffered"; return VLO_ERR_STATE; }
    if (mode == 0 && h->cfg.max_map_points <= 0) { h->err = "handle created with max_map_points = 0"; return VLO_ERR_STATE; }
    const size_t N = (size_t)h->cfg.max_points;
    size_t n_out = 0;
    for (int b = 0; b < n_batches; b++) {
        const vlo_bag_batch &bb = batches[b];
        if (!bb.raw || !bb.offsets || bb.n_scans < 1 || (mode == 0 && !bb.seeds)) return VLO_ERR_INVALID_ARG;
        if (bb.n_scans > HB) { h->err = "a bag batch holds more than max_scans/2 scans"; return VLO_ERR_CAPACITY; }
        for (int s = 0; s < bb.n_scans; s++) {
            int n = bb.offsets[s + 1] - bb.offsets[s];
            if (n < 0) return VLO_ERR_INVALID_ARG;
            if ((size_t)n > N) { h->err = "a scan exceeds max_points"; return VLO_ERR_CAPACITY; }
        }
        size_t total = (size_t)(bb.offsets[bb.n_scans] - bb.offsets[0]);
        if (total * stride + stride > (size_t)HB * N * 4) { h->err = "raw payload of a bag batch exceeds the staging capacity"; return VLO_ERR_CAPACITY; }
        n_out += (size_t)(mode == 0 ? bb.n_scans : bb.n_scans - 1);
    }
    cudaSetDevice(h->cfg.device);
    ScanBatchDev &sb = h->sb;
    if (!h->bag_ctx) h->bag_ctx = new BagCtx();
    BagCtx &ctx = *(BagCtx *)h->bag_ctx;
    BagDrain drain{ ctx };
    ctx.compute_stream = h->stream;
    const size_t par_bytes = sizeof(int) * 2 * (size_t)HB + sizeof(float) * 6 * (size_t)HB;
    if (!ctx.copy_stream) {
        VLO_CUDA(cudaStreamCreateWithFlags(&ctx.copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            VLO_CUDA(cudaEventCreateWithFlags(&ctx.copied[i], cudaEventDisableTiming));
            VLO_CUDA(cudaEventCreateWithFlags(&ctx.raw_free[i], cudaEventDisableTiming));
        }
        VLO_CUDA(cudaMalloc((void **)&ctx.d_par, par_bytes * 2));
    }
    // pinned staging for the whole bag: per batch {offset pairs, slot indices (2 arrays), seeds} + the result records
    const size_t per_batch = sizeof(int) * 4 * (size_t)HB + sizeof(float) * 6 * (size_t)HB;
    const size_t res_off = per_batch * (size_t)n_batches;
    const size_t pin_need = res_off + sizeof(vlo_result) * std::max<size_t>(n_out, 1);
    if (pin_need > ctx.pinned_bytes) {
        if (ctx.pinned) cudaFreeHost(ctx.pinned);
        ctx.pinned = nullptr; ctx.pinned_bytes = 0;
        VLO_CUDA(cudaMallocHost((void **)&ctx.pinned, pin_need + pin_need / 4));
        ctx.pinned_bytes = pin_need + pin_need / 4;
    }
    vlo_result *pres = (vlo_result *)(ctx.pinned + res_off);

    VLO_CUDA(cudaStreamSynchronize(h->stream));
    sb.raw = sb.raw_owned; sb.stride = stride; sb.n_scans = 2 * HB;
    sb.xyz_off[0] = 0; sb.xyz_off[1] = 1; sb.xyz_off[2] = 2;
    h->online_have_last = 0;
    size_t done = 0;
    for (int b = 0; b < n_batches; b++) {
        const vlo_bag_batch &bb = batches[b];
        const int half = b & 1, first = half * HB, n = bb.n_scans;
        char *pp = ctx.pinned + per_batch * (size_t)b;
        int *poff = (int *)pp; int *pidx = poff + 2 * HB; int *pidx2 = pidx + HB; float *pseed = (float *)(pidx2 + HB);
        // this half's raw staging starts at a whole number of points of this stride
        const size_t base_pt = ((size_t)first * N * 4 + stride - 1) / stride;
        for (int s = 0; s < n; s++) { poff[2 * s] = (int)(base_pt + (size_t)(bb.offsets[s] - bb.offsets[0])); poff[2 * s + 1] = bb.offsets[s + 1] - bb.offsets[s]; }
        const size_t total = (size_t)(bb.offsets[n] - bb.offsets[0]);
        // slot indices and seeds of this batch (pinned), uploaded together with the clouds: the compute stream
        // issues no host->device copy of its own, so nothing of batch b ever queues behind batch b+1's clouds on
        // the H2D copy engine
        const int n_res = mode == 0 ? n : n - 1;
        for (int s = 0; s < n_res; s++) { pidx[s] = first + s; pidx2[s] = first + s + 1; }
        const bool have_seeds = bb.seeds != nullptr;
        if (have_seeds && n_res > 0) memcpy(pseed, bb.seeds, sizeof(float) * 6 * (size_t)n_res);
        char *dpar = ctx.d_par + par_bytes * (size_t)half;
        int *d_idx = (int *)dpar, *d_idx2 = d_idx + HB; float *d_seed = (float *)(d_idx2 + HB);
        // ---- copy stream: wait until the kernels that read this half's staging two batches ago are done
        if (b >= 2) VLO_CUDA(cudaStreamWaitEvent(ctx.copy_stream, ctx.raw_free[half], 0));
        VLO_CUDA(cudaMemcpyAsync(dpar, pidx, par_bytes, cudaMemcpyHostToDevice, ctx.copy_stream));
        VLO_CUDA(cudaMemcpyAsync(sb.raw_offset + 2 * first, poff, sizeof(int) * 2 * (size_t)n, cudaMemcpyHostToDevice, ctx.copy_stream));
        // (cudaMemcpyDefault: the clouds may live in host memory -- pinned for the copy to overlap -- or already on the device)
        VLO_CUDA(cudaMemcpyAsync(sb.raw_owned + base_pt * stride, bb.raw + (size_t)bb.offsets[0] * stride, sizeof(float) * total * stride,
                                 cudaMemcpyDefault, ctx.copy_stream));
        VLO_CUDA(cudaEventRecord(ctx.copied[half], ctx.copy_stream));
        // ---- compute stream
        VLO_CUDA(cudaStreamWaitEvent(h->stream, ctx.copied[half], 0));
        sb.scan_first = first; sb.scan_count = n;
        int rc = vlo_launch_organise(h); if (rc) return rc;
        rc = vlo_launch_extract(h); if (rc) return rc;
        h->map_qmax = 0;
        if (mode == 0) {
            rc = vlo_launch_stack_ds(h, first, n); if (rc) return rc;
            rc = vlo_launch_register_map(h, d_idx, n, d_seed); if (rc) return rc;
            VLO_CUDA(cudaMemcpyAsync(pres + done, h->map_result, sizeof(vlo_result) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        } else if (n_res > 0) {
            rc = vlo_build_scan_grids(h, first, n); if (rc) return rc;
            h->grids_valid = 1;
            int *keep_last = h->pair_last, *keep_cur = h->pair_cur;
            h->pair_last = d_idx; h->pair_cur = d_idx2;
            rc = vlo_launch_register_pairs(h, n_res, have_seeds ? d_seed : nullptr, nullptr, -1);
            h->pair_last = keep_last; h->pair_cur = keep_cur;
            if (rc) return rc;
            VLO_CUDA(cudaMemcpyAsync(pres + done, h->pair_result, sizeof(vlo_result) * (size_t)n_res, cudaMemcpyDeviceToHost, h->stream));
        }
        // the staging AND the parameter block of this half are free once this batch's kernels are done
        VLO_CUDA(cudaEventRecord(ctx.raw_free[half], h->stream));
        done += (size_t)n_res;
    }
    h->grids_valid = 0; h->lm.ds_valid = 0;
    VLO_CUDA(cudaStreamSynchronize(ctx.copy_stream));
    int rc = vlo_synchronize(h); if (rc) return rc;
    memcpy(out, pres, sizeof(vlo_result) * n_out);
    int soft = VLO_OK;
    for (size_t k = 0; k < n_out; k++) { vlo_finish_cov_host(&out[k], &h->cfg); if (out[k].status == VLO_SOFT_TOO_FEW_CORR) soft = VLO_SOFT_TOO_FEW_CORR; }
    h->last_n_map = 0; h->last_n_pairs = 0;
    return soft;
}

}  // namespace

void vlo_bag_free(vlo_handle *h)
{
    if (h && h->bag_ctx) { delete (BagCtx *)h->bag_ctx; h->bag_ctx = nullptr; }
}

extern "C" int vlo_bag_register_map(vlo_handle *h, const vlo_bag_batch *batches, int n_batches, int stride_floats, vlo_result *out)
{
    return bag_run(h, batches, n_batches, stride_floats, out, 0);
}

extern "C" int vlo_bag_register_pairs(vlo_handle *h, const vlo_bag_batch *batches, int n_batches, int stride_floats, vlo_result *out)
{
    return bag_run(h, batches, n_batches, stride_floats, out, 1);
}
