// engine.cu -- native frame engine: the three hot methods of emf::EMFusion as ONE host call per frame.
//
//   EMFusion::computeAssociationWeights   reference src/core/EMFusion.cpp:635-670
//   EMFusion::raycast (+ composite)       reference src/core/EMFusion.cpp:726-795
//   EMFusion::integrateDepth              reference src/core/EMFusion.cpp:865-889
//
// The reference drives these with ~2 300 launches, K+1 streams and four host barriers per frame (32 objects).
// Here a frame is: computePoints, association (all volumes, normaliser fused), raycast (all volumes), composite,
// integrate (all volumes, visibility-gated ON THE DEVICE by the composite's counters) -- five launches on one stream,
// no device->host read on the path (visibility counts are copied out asynchronously for the caller's bookkeeping).
//
// The engine owns only per-frame image scratch (points, per-volume raycast / association images, composite outputs,
// counters); the volumes stay caller-owned (emf_volume descriptors).  Host code only: every kernel is reached through
// the level-3 entry points of emf_b200.h.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <new>
#include <vector>

struct emf_engine {
    emf_engine_config cfg;
    int n_vol = 0, has_bg = 0;
    std::vector<emf_volume> vols;
    std::vector<int> ids, gates;
    std::vector<char> force;          // integrate this volume once regardless of its visibility counter (new objects)
    std::vector<int> rects;
    int bg_y0 = 0, bg_y1 = 1 << 30;    // rows of the background's raycast
    // device scratch (one allocation)
    char* pool = nullptr;
    size_t pool_bytes = 0;
    emf_image points{}, norm{}, ray{}, vert{}, nrm{}, seg{}, zero_f{}, zero_f3{}, zero_u8{};
    emf_image partial_target{};   // ptr == nullptr: partial normalisers go to `norm`
    emf_image comp_target[4]{};   // ptr == nullptr: the composite goes to ray / vert / nrm / seg (else: e.g. a slot of an exchange buffer)
    emf_image bg_target[4]{};     // ptr == nullptr: the background's raycast goes to v_ray[0] / v_vert[0] / v_norm[0] / v_mask[0]
    // host-facing frames (emf_engine_submit_host): two slots, three streams
    struct HostSlot {
        float* depth_dev = nullptr; uint8_t* seg_dev = nullptr; float* ray_dev = nullptr;
        uint8_t* seg_host = nullptr; float* ray_host = nullptr;       // pinned
        cudaEvent_t uploaded = nullptr, computed = nullptr, downloaded = nullptr;
    } hs[2];
    cudaStream_t up = nullptr, down = nullptr;
    long long host_count = 0;
    int use_cert = -1;            // ray-space certificate for the background's raycast: -1 = EMF_RAY_CERT decides
    int use_wide = -1;            // four lanes per background ray (EMF_RAY_WIDE): -1 = the environment variable decides
    const int32_t* gate_src = nullptr;   // nullptr: the integrate is gated by vis_count[list position]; else by gate_src[gate_idx[i]]
    std::vector<int> gate_idx;
    std::vector<emf_image> a_img, v_ray, v_vert, v_norm, v_mask;
    int32_t* vis_count = nullptr;
    void* int_ws = nullptr;            // integrate workspace (depth pyramid)
    size_t int_ws_bytes = 0;
    void* ray_ws = nullptr;            // raycast workspace (ray-space certificate)
    size_t ray_ws_bytes = 0;
    int32_t* vis_host = nullptr;       // pinned
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t vis_ready = nullptr;
    cudaStream_t aux = nullptr;        // the association runs here, next to the raycast (both only read the volumes)
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaEvent_t ray_done = nullptr, sched_done = nullptr;   // the raycast's tile sort for the next frame runs on aux, behind the raycast
    bool sched_pending = false;
    cudaStream_t aux2 = nullptr;       // the integrate's preparation (depth pyramid, brick classification)
    cudaEvent_t fork2 = nullptr, join2 = nullptr;
    bool timed_valid = false;
};

namespace {

size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

struct Carver {
    char* base; size_t off;
    emf_image img(int w, int h, size_t elem) {
        emf_image i; i.ptr = base ? base + off : nullptr; i.pitch = (size_t)w * elem; i.width = w; i.height = h;
        off += up256(i.pitch * h);
        return i;
    }
    void* raw(size_t bytes) { void* p = base ? base + off : nullptr; off += up256(bytes); return p; }
};

void carve(emf_engine* e, char* base, size_t* total) {
    Carver c{base, 0};
    const int w = e->cfg.width, h = e->cfg.height, n = e->n_vol;
    e->points = c.img(w, h, 12); e->norm = c.img(w, h, 4);
    e->ray = c.img(w, h, 4); e->vert = c.img(w, h, 12); e->nrm = c.img(w, h, 12); e->seg = c.img(w, h, 1);
    e->zero_f = c.img(w, h, 4); e->zero_f3 = c.img(w, h, 12); e->zero_u8 = c.img(w, h, 1);
    e->a_img.resize(n); e->v_ray.resize(n); e->v_vert.resize(n); e->v_norm.resize(n); e->v_mask.resize(n);
    for (int i = 0; i < n; ++i) {
        e->a_img[i] = c.img(w, h, 4); e->v_ray[i] = c.img(w, h, 4); e->v_vert[i] = c.img(w, h, 12);
        e->v_norm[i] = c.img(w, h, 12); e->v_mask[i] = c.img(w, h, 1);
    }
    e->vis_count = (int32_t*)c.raw(sizeof(int32_t) * EMF_MAX_VOLUMES);
    e->int_ws_bytes = emf_integrate_workspace_bytes(w, h);
    e->int_ws = c.raw(e->int_ws_bytes);
    e->ray_ws_bytes = emf_raycast_workspace_bytes(w, h);
    e->ray_ws = c.raw(e->ray_ws_bytes);
    *total = c.off;
}

}  // namespace

static bool host_slots_ready(emf_engine* e);

extern "C" EMF_API emf_engine* emf_engine_create(const emf_engine_config* cfg) {
    if (!cfg || cfg->width <= 0 || cfg->height <= 0) return nullptr;
    emf_engine* e = new (std::nothrow) emf_engine();
    if (!e) return nullptr;
    e->cfg = *cfg;
    bool ok = cudaMallocHost((void**)&e->vis_host, sizeof(int32_t) * EMF_MAX_VOLUMES) == cudaSuccess;
    for (int k = 0; k < 5 && ok; ++k) ok = cudaEventCreate(&e->ev[k]) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&e->vis_ready, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&e->aux, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&e->fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&e->join, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&e->ray_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&e->sched_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&e->aux2, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&e->fork2, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&e->join2, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && host_slots_ready(e);      // (page-locked allocations are slow: not on the first frame's path)
    if (!ok) { emf_engine_destroy(e); return nullptr; }
    for (int k = 0; k < EMF_MAX_VOLUMES; ++k) e->vis_host[k] = 0;
    return e;
}

extern "C" EMF_API void emf_engine_destroy(emf_engine* e) {
    if (!e) return;
    if (e->pool) cudaFree(e->pool);
    if (e->vis_host) cudaFreeHost(e->vis_host);
    for (int k = 0; k < 5; ++k) if (e->ev[k]) cudaEventDestroy(e->ev[k]);
    if (e->vis_ready) cudaEventDestroy(e->vis_ready);
    if (e->fork) cudaEventDestroy(e->fork);
    if (e->join) cudaEventDestroy(e->join);
    if (e->ray_done) cudaEventDestroy(e->ray_done);
    if (e->sched_done) cudaEventDestroy(e->sched_done);
    for (auto& h : e->hs) {
        if (h.depth_dev) cudaFree(h.depth_dev);
        if (h.seg_dev) cudaFree(h.seg_dev);
        if (h.ray_dev) cudaFree(h.ray_dev);
        if (h.seg_host) cudaFreeHost(h.seg_host);
        if (h.ray_host) cudaFreeHost(h.ray_host);
        if (h.uploaded) cudaEventDestroy(h.uploaded);
        if (h.computed) cudaEventDestroy(h.computed);
        if (h.downloaded) cudaEventDestroy(h.downloaded);
    }
    if (e->up) cudaStreamDestroy(e->up);
    if (e->down) cudaStreamDestroy(e->down);
    if (e->aux) cudaStreamDestroy(e->aux);
    if (e->fork2) cudaEventDestroy(e->fork2);
    if (e->join2) cudaEventDestroy(e->join2);
    if (e->aux2) cudaStreamDestroy(e->aux2);
    delete e;
}

extern "C" EMF_API int emf_engine_set_volumes(emf_engine* e, int n_vol, const emf_volume* vols, int has_background,
                                      emf_stream_t stream) {
    if (!e || n_vol < 0 || (n_vol > 0 && !vols)) return EMF_ERR_INVALID;
    if (n_vol > EMF_MAX_VOLUMES) return EMF_ERR_UNSUPPORTED;
    bool same_ids = n_vol == e->n_vol && (has_background ? 1 : 0) == e->has_bg;
    for (int i = 0; i < n_vol && same_ids; ++i) same_ids = vols[i].id == e->vols[i].id;
    // (a list that changed at equal length -- one object removed, one created -- is re-matched by id like any other change)
    const bool realloc = !same_ids || !e->pool;
    const std::vector<char> old_force = e->force;
    // what the engine held so far: per-volume images survive a change of the volume list (matched by id)
    const std::vector<emf_volume> old_vols = e->vols;
    const int old_has_bg = e->has_bg;
    const std::vector<emf_image> o_a = e->a_img, o_r = e->v_ray, o_v = e->v_vert, o_n = e->v_norm, o_m = e->v_mask;
    const emf_image o_frame[6] = {e->points, e->norm, e->ray, e->vert, e->nrm, e->seg};
    char* old_pool = e->pool;
    e->n_vol = n_vol; e->has_bg = has_background ? 1 : 0;
    e->vols.assign(vols, vols + n_vol);
    e->ids.clear(); e->gates.clear();
    for (int i = 0; i < n_vol; ++i) {
        const bool bg = e->has_bg && i == 0;
        if (!bg) e->ids.push_back(vols[i].id);
        e->gates.push_back(bg ? -1 : (int)e->ids.size() - 1);
    }
    e->force.assign(n_vol, 0);
    for (int i = 0; i < n_vol; ++i)       // pending force flags follow their volume
        for (int k = 0; k < (int)old_vols.size() && k < (int)old_force.size(); ++k)
            if (old_vols[k].id == vols[i].id && ((old_has_bg && k == 0) == (e->has_bg && i == 0))) { e->force[i] = old_force[k]; break; }
    e->rects.assign((size_t)4 * (n_vol > 0 ? n_vol : 1), 0);
    cudaStream_t s = (cudaStream_t)stream;
    if (realloc) {
        size_t total = 0;
        carve(e, nullptr, &total);
        char* pool = nullptr;
        if (cudaMalloc((void**)&pool, total) != cudaSuccess) { e->pool = old_pool; e->vols = old_vols; e->n_vol = (int)old_vols.size(); e->has_bg = old_has_bg; carve(e, old_pool, &total); return EMF_ERR_CUDA; }
        e->pool = pool; e->pool_bytes = total;
        carve(e, e->pool, &total);
        cudaMemsetAsync(e->pool, 0, total, s);
        const size_t w = e->cfg.width, h = e->cfg.height;
        if (old_pool) {
            const emf_image n_frame[6] = {e->points, e->norm, e->ray, e->vert, e->nrm, e->seg};
            for (int k = 0; k < 6; ++k) cudaMemcpyAsync(n_frame[k].ptr, o_frame[k].ptr, n_frame[k].pitch * h, cudaMemcpyDeviceToDevice, s);
        }
        for (int i = 0; i < n_vol; ++i) {
            const bool bg = e->has_bg && i == 0;
            int j = -1;
            for (int k = 0; k < (int)old_vols.size() && old_pool; ++k)
                if ((old_has_bg && k == 0) == bg && old_vols[k].id == vols[i].id) { j = k; break; }
            if (j >= 0) {
                cudaMemcpyAsync(e->a_img[i].ptr, o_a[j].ptr, w * h * 4, cudaMemcpyDeviceToDevice, s);
                cudaMemcpyAsync(e->v_ray[i].ptr, o_r[j].ptr, w * h * 4, cudaMemcpyDeviceToDevice, s);
                cudaMemcpyAsync(e->v_vert[i].ptr, o_v[j].ptr, w * h * 12, cudaMemcpyDeviceToDevice, s);
                cudaMemcpyAsync(e->v_norm[i].ptr, o_n[j].ptr, w * h * 12, cudaMemcpyDeviceToDevice, s);
                cudaMemcpyAsync(e->v_mask[i].ptr, o_m[j].ptr, w * h, cudaMemcpyDeviceToDevice, s);
            } else {
                // a new volume's association image starts at 1 (reference src/core/EMFusion.cpp:55,916)
                emf_fill_image_f32(&e->a_img[i], 1.0f, stream);
            }
        }
        if (old_pool) { cudaStreamSynchronize(s); cudaFree(old_pool); }
    } else {
        carve(e, e->pool, &e->pool_bytes);
    }
    return emfb::launch_status();
}

extern "C" EMF_API int emf_engine_frame(emf_engine* e, const emf_image* depth, const emf_pose* T_co, const emf_pose* T_oc,
                                unsigned flags, emf_stream_t stream) {
    if (!e || !e->pool) return EMF_ERR_INVALID;
    const int n = e->n_vol, w = e->cfg.width, h = e->cfg.height;
    cudaStream_t s = (cudaStream_t)stream;
    const bool timed = (flags & EMF_FRAME_TIMED) != 0;
    int rc = EMF_OK;
    bool overlap = false;
    e->timed_valid = false;
    if (flags & EMF_FRAME_POINTS) {
        if (!emfb::image_ok(depth, 4) || depth->width != w || depth->height != h) return EMF_ERR_INVALID;
        rc = emf_compute_points(depth, &e->points, e->cfg.K, stream);
        if (rc != EMF_OK) return rc;
    }
    if (timed) cudaEventRecord(e->ev[0], s);
    // what the integrate needs from the depth image and the poses alone (depth pyramid, brick classification) runs on a side
    // stream next to the association and the raycast
    bool prepared = false;
    if ((flags & EMF_FRAME_INTEGRATE) && (flags & EMF_FRAME_RAYCAST) && n > 0 && T_oc && emfb::image_ok(depth, 4) && !timed) {
        cudaEventRecord(e->fork2, s);
        cudaStreamWaitEvent(e->aux2, e->fork2, 0);
        rc = emf_integrate_volumes_phase(n, e->vols.data(), T_oc, e->cfg.K, depth, e->a_img.data(), e->cfg.params.max_tsdf_weight,
                                         nullptr, nullptr, 0, nullptr, e->int_ws, e->int_ws_bytes, 1, (emf_stream_t)e->aux2);
        if (rc != EMF_OK) return rc;
        cudaEventRecord(e->join2, e->aux2);
        prepared = true;
    }
    if ((flags & (EMF_FRAME_ASSOC | EMF_FRAME_ASSOC_PARTIAL | EMF_FRAME_ASSOC_PARTIAL_NOBG)) && n > 0) {
        if (!T_co) return EMF_ERR_INVALID;
        const int mode = (flags & EMF_FRAME_ASSOC_PARTIAL_NOBG) ? (e->has_bg ? 3 : 1) : ((flags & EMF_FRAME_ASSOC_PARTIAL) ? 1 : 0);
        // association and raycast of one call only read the volumes and write disjoint images: when both are asked for
        // (and no stage timing is wanted) the association runs on a side stream and is joined before the integrate
        overlap = mode == 0 && (flags & EMF_FRAME_RAYCAST) && !timed;
        if (overlap) {
            cudaEventRecord(e->fork, s);
            cudaStreamWaitEvent(e->aux, e->fork, 0);
        }
        rc = emf_assoc_weights(n, e->vols.data(), T_co, &e->points, &e->cfg.params, e->a_img.data(), mode,
                               (mode != 0 && e->partial_target.ptr) ? &e->partial_target : &e->norm,
                               overlap ? (emf_stream_t)e->aux : stream);
        if (rc != EMF_OK) return rc;
        if (overlap) cudaEventRecord(e->join, e->aux);
    } else if ((flags & (EMF_FRAME_ASSOC_PARTIAL | EMF_FRAME_ASSOC_PARTIAL_NOBG)) && n == 0) {
        const emf_image& tgt = e->partial_target.ptr ? e->partial_target : e->norm;
        cudaMemsetAsync(tgt.ptr, 0, tgt.pitch * h, s);
    }
    if ((flags & EMF_FRAME_NORMALISE) && n > 0) {
        rc = emf_assoc_normalise(n, e->a_img.data(), &e->norm, stream);
        if (rc != EMF_OK) return rc;
    }
    if (timed) cudaEventRecord(e->ev[1], s);
    if ((flags & EMF_FRAME_RAYCAST) && n > 0) {
        if (!T_co) return EMF_ERR_INVALID;
        for (int i = 0; i < n; ++i) {
            rc = emf_volume_screen_rect(e->vols[i].res, e->vols[i].voxel_size, &T_co[i], e->cfg.K, w, h, &e->rects[4 * i]);
            if (rc != EMF_OK) return rc;
        }
        if (e->has_bg && (e->rects[0] > 0 || e->rects[1] > 0 || e->rects[2] < w || e->rects[3] < h))
            // the composite reads the background's hit mask over the whole frame: pixels its box cannot cover are "no hit"
            cudaMemsetAsync(e->bg_target[3].ptr ? e->bg_target[3].ptr : e->v_mask[0].ptr, 0, e->v_mask[0].pitch * h, s);
        if (e->has_bg) {   // band of rows of the background (multi-GPU, replicated background)
            e->rects[1] = std::max(e->rects[1], e->bg_y0);
            e->rects[3] = std::max(e->rects[1], std::min(e->rects[3], e->bg_y1));
        }
        // the ray-space certificate (emf_raycast_volumes_ws) is opt-in: EMF_RAY_CERT=1 or emf_engine_set_option (see
        // DESIGN.md for the measurements: it pays when a GPU traces a band of the frame, not the whole of it)
        static const bool env_cert = [] { const char* v = getenv("EMF_RAY_CERT"); return v && v[0] == '1'; }();
        const bool use_cert = e->use_cert < 0 ? env_cert : e->use_cert != 0;
        std::vector<emf_image> r_ray(e->v_ray), r_vert(e->v_vert), r_norm(e->v_norm), r_mask(e->v_mask);
        if (e->has_bg && e->bg_target[0].ptr) { r_ray[0] = e->bg_target[0]; r_vert[0] = e->bg_target[1]; r_norm[0] = e->bg_target[2]; r_mask[0] = e->bg_target[3]; }
        static const bool env_wide = [] { const char* v = getenv("EMF_RAY_WIDE"); return v && v[0] == '1'; }();
        const bool use_wide = e->use_wide < 0 ? env_wide : e->use_wide != 0;
        static const bool env_sched = [] { const char* v = getenv("EMF_RAY_SCHED"); return !(v && v[0] == '0'); }();
        const bool sched = env_sched && !use_cert && !use_wide;
        if (e->sched_pending) { cudaStreamWaitEvent(s, e->sched_done, 0); e->sched_pending = false; }
        rc = emf_raycast_volumes_opt(n, e->vols.data(), T_co, e->cfg.K, e->rects.data(), r_ray.data(), r_vert.data(),
                                     r_norm.data(), r_mask.data(), nullptr, e->ray_ws, e->ray_ws_bytes,
                                     (use_cert ? EMF_RAY_CERTIFICATE : 0u) | (sched ? (EMF_RAY_SCHEDULE | EMF_RAY_SCHEDULE_DEFER) : 0u) |
                                         (use_wide && !use_cert ? EMF_RAY_WIDE : 0u), stream);
        if (rc != EMF_OK) return rc;
        if (sched) {      // the sort of the next frame's tile order: on the side stream, under the composite / integrate
            cudaEventRecord(e->ray_done, s);
            cudaStreamWaitEvent(e->aux, e->ray_done, 0);
            rc = emf_raycast_schedule_update(w, h, e->rects.data(), e->ray_ws, e->ray_ws_bytes, (emf_stream_t)e->aux);
            if (rc != EMF_OK) return rc;
            cudaEventRecord(e->sched_done, e->aux);
            e->sched_pending = true;
        }
    }
    if (flags & (EMF_FRAME_COMPOSITE | EMF_FRAME_COMPOSITE_NOBG)) {
        // (force flags -- objects created since the last integrate, reference src/core/EMFusion.cpp:550,918: createObj puts
        //  them into vis_objs after the raycast -- stay set until an integrate consumes them: a frame driven through one
        //  emf_engine_frame call creates its objects before the raycast, and the composite must not drop them)
        const int o0 = e->has_bg ? 1 : 0, n_obj = n - o0;
        const bool with_bg = e->has_bg && !(flags & EMF_FRAME_COMPOSITE_NOBG);
        const emf_image* bg_ray = with_bg ? &e->v_ray[0] : &e->zero_f;
        const emf_image* bg_vert = with_bg ? &e->v_vert[0] : &e->zero_f3;
        const emf_image* bg_norm = with_bg ? &e->v_norm[0] : &e->zero_f3;
        const emf_image* bg_mask = with_bg ? &e->v_mask[0] : &e->zero_u8;
        rc = emf_raycast_composite(n_obj, e->ids.data(), e->rects.data() + 4 * o0, e->v_ray.data() + o0, e->v_vert.data() + o0,
                                   e->v_norm.data() + o0, e->v_mask.data() + o0, bg_ray, bg_vert, bg_norm, bg_mask,
                                   e->cfg.boundary, e->comp_target[0].ptr ? &e->comp_target[0] : &e->ray,
                                   e->comp_target[0].ptr ? &e->comp_target[1] : &e->vert, e->comp_target[0].ptr ? &e->comp_target[2] : &e->nrm,
                                   e->comp_target[0].ptr ? &e->comp_target[3] : &e->seg, e->vis_count, stream);
        if (rc != EMF_OK) return rc;
        if (n_obj > 0) {
            cudaMemcpyAsync(e->vis_host, e->vis_count, sizeof(int32_t) * n_obj, cudaMemcpyDeviceToHost, s);
            cudaEventRecord(e->vis_ready, s);
        }
    }
    if (overlap) cudaStreamWaitEvent(s, e->join, 0);   // (also when no integrate follows: the caller sees one stream)
    if (timed) cudaEventRecord(e->ev[2], s);
    if ((flags & (EMF_FRAME_INTEGRATE | EMF_FRAME_INTEGRATE_BG | EMF_FRAME_INTEGRATE_OBJ)) && n > 0) {
        if (!T_oc || !emfb::image_ok(depth, 4)) return EMF_ERR_INVALID;
        // the whole list, or only the background (which no visibility counter gates: it need not wait for the composite), or
        // only the objects
        int i0 = 0, i1 = n;
        if (!(flags & EMF_FRAME_INTEGRATE)) {
            if ((flags & EMF_FRAME_INTEGRATE_BG) && !(flags & EMF_FRAME_INTEGRATE_OBJ)) i1 = e->has_bg ? 1 : 0;
            else if ((flags & EMF_FRAME_INTEGRATE_OBJ) && !(flags & EMF_FRAME_INTEGRATE_BG)) i0 = e->has_bg ? 1 : 0;
        }
        if (i1 > i0) {
            const emf_image* assoc = e->a_img.data() + i0;
            const bool gate = (flags & EMF_FRAME_INTEGRATE_ALL) == 0;
            std::vector<int> g(e->gates.begin() + i0, e->gates.begin() + i1);
            const int32_t* counts = e->vis_count;
            if (e->gate_src && (int)e->gate_idx.size() == n) {      // (multi-GPU: the merged frame's counters, by global list position)
                counts = e->gate_src;
                for (int i = i0; i < i1; ++i) if (g[i - i0] >= 0) g[i - i0] = e->gate_idx[i];
            }
            for (int i = i0; i < i1; ++i) if (e->force[i]) { g[i - i0] = -1; e->force[i] = 0; }
            const bool prep = prepared && i0 == 0 && i1 == n;
            if (prep) cudaStreamWaitEvent(s, e->join2, 0);
            rc = emf_integrate_volumes_phase(i1 - i0, e->vols.data() + i0, T_oc + i0, e->cfg.K, depth, assoc, e->cfg.params.max_tsdf_weight,
                                             gate ? counts : nullptr, gate ? g.data() : nullptr,
                                             e->cfg.visibility_thresh, nullptr, e->int_ws, e->int_ws_bytes, prep ? 2 : 0, stream);
            if (rc != EMF_OK) return rc;
            rc = emf_update_brick_maps(i1 - i0, e->vols.data() + i0, stream);
            if (rc != EMF_OK) return rc;
        }
    }
    if (timed) { cudaEventRecord(e->ev[3], s); e->timed_valid = true; }
    return emfb::launch_status();
}

extern "C" EMF_API int emf_engine_set_partial_norm_target(emf_engine* e, const emf_image* target) {
    if (!e) return EMF_ERR_INVALID;
    if (!target) { e->partial_target = emf_image{}; return EMF_OK; }
    if (!emfb::image_ok(target, 4) || target->width != e->cfg.width || target->height != e->cfg.height) return EMF_ERR_INVALID;
    e->partial_target = *target;
    return EMF_OK;
}

// ---------------------------------------------------------------------------------------------
// Host-facing frame: what emf::EMFusion::processFrame does between `depth_raw.upload` (reference src/core/EMFusion.cpp:72)
// and the host-side consumers of the raycast (getLastMasks / the renderers, :131-200), with host buffers at both ends.
// The reference uploads, processes and downloads strictly one after the other, blocking the host on each.  Here the upload
// of frame n + 1 and the download of frame n run on their own streams under the kernels of the neighbouring frames; the
// host only waits when it asks for a result.  Every frame's depth crosses PCIe once, every result is read back once.
// ---------------------------------------------------------------------------------------------
static bool host_slots_ready(emf_engine* e) {
    if (e->up) return true;
    const size_t px = (size_t)e->cfg.width * e->cfg.height;
    bool ok = cudaStreamCreateWithFlags(&e->up, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&e->down, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& h : e->hs) {
        ok = ok && cudaMalloc((void**)&h.depth_dev, px * 4) == cudaSuccess && cudaMalloc((void**)&h.seg_dev, px) == cudaSuccess &&
             cudaMalloc((void**)&h.ray_dev, px * 4) == cudaSuccess && cudaMallocHost((void**)&h.seg_host, px) == cudaSuccess &&
             cudaMallocHost((void**)&h.ray_host, px * 4) == cudaSuccess &&
             cudaEventCreateWithFlags(&h.uploaded, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&h.computed, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&h.downloaded, cudaEventDisableTiming) == cudaSuccess;
    }
    return ok;
}

extern "C" EMF_API int emf_engine_submit_host(emf_engine* e, const float* depth_host, const emf_pose* T_co, const emf_pose* T_oc,
                                      unsigned flags, int download, emf_stream_t stream, long long* ticket_out) {
    if (!e || !e->pool || !depth_host) return EMF_ERR_INVALID;
    if (!host_slots_ready(e)) return EMF_ERR_CUDA;
    const int w = e->cfg.width, h = e->cfg.height;
    const size_t px = (size_t)w * h;
    cudaStream_t s = (cudaStream_t)stream;
    const long long k = e->host_count;
    emf_engine::HostSlot& H = e->hs[k & 1];
    if (k >= 2) cudaStreamWaitEvent(e->up, H.computed, 0);          // the frame that last used this depth slot is done with it
    cudaMemcpyAsync(H.depth_dev, depth_host, px * 4, cudaMemcpyHostToDevice, e->up);
    cudaEventRecord(H.uploaded, e->up);
    cudaStreamWaitEvent(s, H.uploaded, 0);
    if (k >= 2) cudaStreamWaitEvent(s, H.downloaded, 0);            // ... and its staged result has left the device
    emf_image depth{H.depth_dev, (size_t)w * 4, w, h};
    const int rc = emf_engine_frame(e, &depth, T_co, T_oc, flags, stream);
    if (rc != EMF_OK) return rc;
    if (download) {   // stage the composite on the compute stream so that the next frame may overwrite the engine's images
        cudaMemcpyAsync(H.seg_dev, e->seg.ptr, px, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(H.ray_dev, e->ray.ptr, px * 4, cudaMemcpyDeviceToDevice, s);
    }
    cudaEventRecord(H.computed, s);
    cudaStreamWaitEvent(e->down, H.computed, 0);
    if (download) {
        cudaMemcpyAsync(H.seg_host, H.seg_dev, px, cudaMemcpyDeviceToHost, e->down);
        cudaMemcpyAsync(H.ray_host, H.ray_dev, px * 4, cudaMemcpyDeviceToHost, e->down);
    }
    cudaEventRecord(H.downloaded, e->down);
    if (ticket_out) *ticket_out = k;
    e->host_count = k + 1;
    return emfb::launch_status();
}

extern "C" EMF_API int emf_engine_result_host(emf_engine* e, long long ticket, const uint8_t** seg_host, const float** ray_host) {
    if (!e || !e->up || ticket < e->host_count - 2 || ticket >= e->host_count) return EMF_ERR_INVALID;
    emf_engine::HostSlot& H = e->hs[ticket & 1];
    if (cudaEventSynchronize(H.downloaded) != cudaSuccess) return EMF_ERR_CUDA;
    if (seg_host) *seg_host = H.seg_host;
    if (ray_host) *ray_host = H.ray_host;
    return EMF_OK;
}

extern "C" EMF_API int emf_engine_depth_slot(emf_engine* e, long long ticket, emf_image* depth_out) {
    if (!e || !e->up || !depth_out || ticket < e->host_count - 2 || ticket >= e->host_count) return EMF_ERR_INVALID;
    *depth_out = emf_image{e->hs[ticket & 1].depth_dev, (size_t)e->cfg.width * 4, e->cfg.width, e->cfg.height};
    return EMF_OK;
}

extern "C" EMF_API int emf_engine_set_composite_target(emf_engine* e, const emf_image* target4) {
    if (!e) return EMF_ERR_INVALID;
    if (!target4) { for (int k = 0; k < 4; ++k) e->comp_target[k] = emf_image{}; return EMF_OK; }
    const size_t el[4] = {4, 12, 12, 1};
    for (int k = 0; k < 4; ++k)
        if (!emfb::image_ok(&target4[k], el[k]) || target4[k].width != e->cfg.width || target4[k].height != e->cfg.height) return EMF_ERR_INVALID;
    for (int k = 0; k < 4; ++k) e->comp_target[k] = target4[k];
    return EMF_OK;
}

extern "C" EMF_API int emf_engine_set_background_target(emf_engine* e, const emf_image* target4) {
    if (!e) return EMF_ERR_INVALID;
    if (!target4) { for (int k = 0; k < 4; ++k) e->bg_target[k] = emf_image{}; return EMF_OK; }
    const size_t el[4] = {4, 12, 12, 1};
    for (int k = 0; k < 4; ++k)
        if (!emfb::image_ok(&target4[k], el[k]) || target4[k].width != e->cfg.width || target4[k].height != e->cfg.height) return EMF_ERR_INVALID;
    for (int k = 0; k < 4; ++k) e->bg_target[k] = target4[k];
    return EMF_OK;
}

extern "C" EMF_API int emf_engine_set_gate_source(emf_engine* e, const int32_t* counts, const int* index, int n) {
    if (!e) return EMF_ERR_INVALID;
    if (!counts) { e->gate_src = nullptr; e->gate_idx.clear(); return EMF_OK; }
    if (!index || n != e->n_vol) return EMF_ERR_INVALID;
    e->gate_src = counts;
    e->gate_idx.assign(index, index + n);
    return EMF_OK;
}

extern "C" EMF_API int emf_engine_set_option(emf_engine* e, int option, int value) {
    if (!e) return EMF_ERR_INVALID;
    switch (option) {
    case EMF_OPT_RAY_CERTIFICATE: e->use_cert = value; return EMF_OK;
    case EMF_OPT_RAY_WIDE: e->use_wide = value; return EMF_OK;
    default: return EMF_ERR_INVALID;
    }
}

extern "C" EMF_API int emf_engine_normalise_from_parts(emf_engine* e, int n_parts, const float* const* parts, const uint32_t* flags,
                                                       uint32_t value, uint32_t* err, double timeout_s, emf_stream_t stream) {
    if (!e || !e->pool) return EMF_ERR_INVALID;
    return emf_assoc_normalise_parts(e->n_vol, e->a_img.data(), n_parts, parts, &e->norm, flags, value, err, timeout_s, stream);
}

extern "C" EMF_API int emf_engine_stage_ms(emf_engine* e, float ms[3]) {
    if (!e || !ms || !e->timed_valid) return EMF_ERR_INVALID;
    if (cudaEventSynchronize(e->ev[3]) != cudaSuccess) return EMF_ERR_CUDA;
    for (int k = 0; k < 3; ++k)
        if (cudaEventElapsedTime(&ms[k], e->ev[k], e->ev[k + 1]) != cudaSuccess) return EMF_ERR_CUDA;
    return EMF_OK;
}

extern "C" EMF_API int emf_engine_image(emf_engine* e, int what, int index, emf_image* out) {
    if (!e || !out || !e->pool) return EMF_ERR_INVALID;
    const bool vi = index >= 0 && index < e->n_vol;
    switch (what) {
    case EMF_IMG_POINTS: *out = e->points; return EMF_OK;
    case EMF_IMG_NORM: *out = e->norm; return EMF_OK;
    case EMF_IMG_RAY: *out = e->ray; return EMF_OK;
    case EMF_IMG_VERT: *out = e->vert; return EMF_OK;
    case EMF_IMG_NORMALS: *out = e->nrm; return EMF_OK;
    case EMF_IMG_SEG: *out = e->seg; return EMF_OK;
    case EMF_IMG_VOL_ASSOC: if (!vi) return EMF_ERR_INVALID; *out = e->a_img[index]; return EMF_OK;
    case EMF_IMG_VOL_RAY: if (!vi) return EMF_ERR_INVALID; *out = e->v_ray[index]; return EMF_OK;
    case EMF_IMG_VOL_VERT: if (!vi) return EMF_ERR_INVALID; *out = e->v_vert[index]; return EMF_OK;
    case EMF_IMG_VOL_NORMALS: if (!vi) return EMF_ERR_INVALID; *out = e->v_norm[index]; return EMF_OK;
    case EMF_IMG_VOL_MASK: if (!vi) return EMF_ERR_INVALID; *out = e->v_mask[index]; return EMF_OK;
    default: return EMF_ERR_INVALID;
    }
}

extern "C" EMF_API int32_t* emf_engine_vis_counts_device(emf_engine* e) { return e ? e->vis_count : nullptr; }

extern "C" EMF_API int emf_engine_vis_counts(emf_engine* e, int32_t* counts_out, int n) {
    if (!e || !counts_out || n < 0 || n > EMF_MAX_VOLUMES) return EMF_ERR_INVALID;
    if (cudaEventSynchronize(e->vis_ready) != cudaSuccess) return EMF_ERR_CUDA;
    for (int k = 0; k < n; ++k) counts_out[k] = e->vis_host[k];
    return EMF_OK;
}

extern "C" EMF_API int emf_engine_force_integrate(emf_engine* e, int vol_index) {
    if (!e || vol_index < 0 || vol_index >= e->n_vol) return EMF_ERR_INVALID;
    e->force[vol_index] = 1;
    return EMF_OK;
}

extern "C" EMF_API int emf_engine_set_background_rows(emf_engine* e, int y0, int y1) {
    if (!e || y0 < 0 || y1 < y0) return EMF_ERR_INVALID;
    e->bg_y0 = y0; e->bg_y1 = y1;
    return EMF_OK;
}
