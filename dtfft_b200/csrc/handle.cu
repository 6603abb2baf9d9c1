// ReshapeHandle (see handle.h for the reference citations).
#include "handle.h"

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "errors.h"

namespace dtfftb {

void ReshapeHandle::destroy() {
    peer_pieces_[0].clear();
    peer_pieces_[1].clear();
    dma_pack_.reset();
    dma_self_.reset();
    dma_subs_.clear();
    dma_ = dma_standalone_ = false;
    dma_bases_ = nullptr;
    for (int i = 0; i < kCopyStreams; ++i) {
        if (copy_streams_[i]) cudaStreamDestroy(copy_streams_[i]);
        if (copies_done_[i]) cudaEventDestroy(copies_done_[i]);
        copy_streams_[i] = nullptr, copies_done_[i] = nullptr;
    }
    for (cudaEvent_t e : pack_done_) cudaEventDestroy(e);
    pack_done_.clear();
    fused_.clear();
    fused_chunks_.clear();
    if (ctx_.peers)
        for (auto& kv : maps_) ctx_.peers->release(&kv.second.opened);
    maps_.clear();
    fused_boxes_.clear();
    nccl_.reset();
    pack_.reset();
    unpack_.reset();
    created_ = false;
}

int ReshapeHandle::create(const HandleContext& ctx, int ttype, int rtype, int comm_id, const std::vector<int>& members,
                          int me, const std::vector<Pencil>& send_by_member, const std::vector<Pencil>& recv_by_member,
                          int64_t base_storage, int backend) {
    destroy();
    ctx_ = ctx;
    is_transpose_ = ttype != 0;
    comm_id_ = comm_id;
    members_ = members;
    me_ = me;
    es_ = base_storage;
    backend_ = backend;
    const int P = (int)members.size();
    has_exchange_ = P > 1;
    aux_bytes_ = 0;
    const Pencil& send = send_by_member[(size_t)me];
    const Pencil& recv = recv_by_member[(size_t)me];
    const int ndims = send.ndims;
    send_elems_ = send.size();
    recv_elems_ = recv.size();
    send_ = send;
    recv_by_member_ = recv_by_member;
    send_by_member_ = send_by_member;
    int rc;

    // element counts that cross the NVLink fabric (everything not addressed to myself)
    {
        const RankLayout src = layout_of(send);
        long long self = 0;
        bool tr = false;
        Box b = intersect_box(src, layout_of(recv), &tr);
        if (!b.empty()) self = b.volume();
        remote_elems_ = send_elems_ - self;
    }

    if (!has_exchange_) {  // reshape_handle_generic.F90:246-253
        pack_.reset(new Kernel);
        geo_ = HandleGeometry{};
        geo_.ttype = ttype, geo_.rtype = rtype, geo_.ndims = ndims;
        int kt = K_COPY;
        if (is_transpose_) {
            const bool fwd = ttype == T_X_TO_Y || ttype == T_Y_TO_Z || ttype == T_Z_TO_X;
            kt = fwd ? K_PERMUTE_FORWARD : K_PERMUTE_BACKWARD;
        }
        geo_.pack_kernel = kt;
        rc = pack_->create(ndims, send.counts, kt, es_, nullptr, 0, ctx_.effort, false);
        if (rc) return rc;
        launches_ = kt == K_COPY ? 0 : 1;
        created_ = true;
        return DTFFT_SUCCESS;
    }

    if (backend_ == BACKEND_NVLINK_FUSED) {
        if (!ctx_.peers || !ctx_.peers->available()) return DTFFT_ERROR_INVALID_BACKEND;
        const RankLayout src = layout_of(send);
        fused_boxes_.assign((size_t)P, Box{});
        bool any_t = false, any_r = false;
        for (int i = 0; i < P; ++i) {
            bool tr = false;
            Box b = intersect_box(src, layout_of(recv_by_member[(size_t)i]), &tr);
            fused_boxes_[(size_t)i] = b;
            if (!b.empty()) (tr ? any_t : any_r) = true;
        }
        if (any_t && any_r) return DTFFTB_ERROR_INTERNAL;
        fused_family_ = any_r ? FAM_R : FAM_T;
        // ---- copy-engine form (handle.h): every member must take the same decision, so it is taken from what
        // every member knows -- all pencils of the group
        {
            bool all_ok = true;
            long long max_block_bytes = 0;
            for (int r = 0; r < P && all_ok; ++r) {
                for (int i = 0; i < P && all_ok; ++i) {
                    if (i == r) continue;
                    const int nsub = dma_nsub(send_by_member[(size_t)r], recv_by_member[(size_t)i], es_, P);
                    long long bytes = 0;
                    for (int q = 0; q < nsub && all_ok; ++q) {
                        bool tr = false;
                        const Box b = block_box(send_by_member[(size_t)r], recv_by_member[(size_t)i], q, nsub, &tr);
                        if (b.empty()) continue;
                        all_ok = dma_block(b, tr, 0).ok;
                        bytes += b.volume() * es_;
                    }
                    max_block_bytes = std::max(max_block_bytes, bytes);
                }
            }
            const char* e = getenv("DTFFTB_FUSED_MODE");
            const bool force_dma = e && (e[0] == 'd' || e[0] == 'D');
            const bool force_store = e && (e[0] == 's' || e[0] == 'S');
            // Where the copy-engine form pays (B200, 512^3 c128 cycle, profiles/): on its own it is slower than the
            // direct-store kernel (every copy costs ~13 us of its own: 0.835 vs 0.787 ms per exchange at 2 GPUs, 0.444 vs
            // 0.370 ms at 8), but pipelined peer by peer with the local transposition next to it the cycle drops from
            // 2.20 to 1.91 ms at 2 GPUs and from 1.51 to 1.37 ms at 4; at 8 GPUs the local transpositions are too short to
            // pay for the copies (0.99 vs 0.89 ms).  So by default it is kept for the pair pipeline of groups of up to 4
            // with large blocks, and a lone transposition runs the direct-store kernel.  DTFFTB_FUSED_MODE = dma forces it
            // everywhere it is possible, = store switches it off.
            int max_group = 4;
            if (const char* g = getenv("DTFFTB_DMA_MAX_GROUP")) max_group = atoi(g);
            dma_ = all_ok && !force_store && (force_dma || (max_block_bytes >= (8ll << 20) && P <= max_group));
            dma_standalone_ = dma_ && force_dma;
            if (dma_) {
                dma_subs_.assign((size_t)P, std::vector<DmaSub>());
                std::vector<Box> packs, selfs((size_t)P);
                selfs[(size_t)me] = fused_boxes_[(size_t)me];
                long long off = 0;
                for (int i = 0; i < P; ++i) {
                    if (i == me) continue;
                    const int nsub = dma_nsub(send, recv_by_member[(size_t)i], es_, P);
                    for (int q = 0; q < nsub; ++q) {
                        bool tr = false;
                        const Box b = block_box(send, recv_by_member[(size_t)i], q, nsub, &tr);
                        DmaSub sub;
                        if (!b.empty()) {
                            sub.blk = dma_block(b, tr, off);
                            sub.pack_index = (int)packs.size();
                            packs.push_back(sub.blk.pack);
                            off += b.volume();
                        }
                        dma_subs_[(size_t)i].push_back(sub);
                    }
                }
                dma_pack_.reset(new Kernel);
                if (!packs.empty()) {
                    rc = dma_pack_->create_boxes(fused_family_, es_, packs);
                    if (rc) return rc;
                }
                dma_self_.reset(new Kernel);
                rc = dma_self_->create_boxes(fused_family_, es_, selfs);
                if (rc) return rc;
                aux_bytes_ = off * es_;  // staging of the blocks that leave the GPU
            }
        }
        geo_ = HandleGeometry{};
        geo_.ttype = ttype, geo_.rtype = rtype, geo_.ndims = ndims;
        geo_.comm_size = P, geo_.comm_rank = me, geo_.members = members;
        geo_.has_exchange = true, geo_.is_fused = true;
        launches_ = dma_standalone_ ? 2 + P : 3;  // barrier + (packs + self | fused kernel) + barrier
        created_ = true;
        return DTFFT_SUCCESS;
    }

    if (backend_ != BACKEND_NCCL && backend_ != BACKEND_NCCL_PIPELINED) return DTFFT_ERROR_INVALID_BACKEND;
    if (!ctx_.nccl) return DTFFTB_ERROR_INTERNAL;
    const bool pipelined = backend_is_pipelined(backend_);
    pack_.reset(new Kernel);
    unpack_.reset(new Kernel);
    nccl_.reset(new NcclBackend);

    if (is_transpose_) {
        // the reference's neighbor_data tables are int32 element offsets: a pencil beyond 2^31 - 1
        // elements is an INTERNAL_ERROR there (check_if_overflow, :159-172).  Only this path keeps
        // the limit (NVLINK_FUSED, brick reshapes and single-rank plans use 64-bit boxes).
        for (int i = 0; i < P; ++i)
            if (send_by_member[(size_t)i].size() > INT32_MAX || recv_by_member[(size_t)i].size() > INT32_MAX)
                return DTFFTB_ERROR_INTERNAL;
        geo_ = transpose_geometry(ttype, send_by_member, recv_by_member, me, members, pipelined, false);
        rc = pack_->create(ndims, geo_.send_dims, geo_.pack_kernel, es_, geo_.send_nd.data(), P, ctx_.effort, false);
        if (rc) return rc;
        rc = unpack_->create(ndims, geo_.recv_dims, geo_.unpack_kernel, es_, geo_.recv_nd.data(), P, ctx_.effort, false);
        if (rc) return rc;
    } else {
        // brick <-> pencil reshape: same axis order on both sides; block (me -> i) is the
        // global-index intersection, carried in a contiguous slot in destination order
        geo_ = HandleGeometry{};
        geo_.ttype = 0, geo_.rtype = rtype, geo_.ndims = ndims;
        geo_.comm_size = P, geo_.comm_rank = me, geo_.members = members;
        geo_.has_exchange = true, geo_.is_pipelined = pipelined;
        geo_.pack_kernel = K_PACK;  // :438
        geo_.unpack_kernel = pipelined ? K_UNPACK_PIPELINED : K_UNPACK;
        for (int j = 0; j < 3; ++j) geo_.send_dims[j] = j < ndims ? send.counts[j] : 1, geo_.recv_dims[j] = j < ndims ? recv.counts[j] : 1;
        const ReshapeGeometry rg = reshape_geometry(rtype, send_by_member, recv_by_member, me);
        geo_.send_counts = rg.send_counts, geo_.send_displs = rg.send_displs;
        geo_.recv_counts = rg.recv_counts, geo_.recv_displs = rg.recv_displs;
        geo_.reshape_strat = rg.reshape_strat;
        // pack-free / unpack-free shortcuts (:261-266, 479-484); DTFFTB_RESHAPE_SHORTCUTS=0 keeps
        // the three-step schedule (every rank must set it alike: aux sizes change with it)
        const char* sc = getenv("DTFFTB_RESHAPE_SHORTCUTS");
        const bool shortcuts = !(sc && sc[0] == '0') && !ctx_.no_shortcuts;
        geo_.is_pack_free = shortcuts && rg.is_pack_free;
        geo_.is_unpack_free = shortcuts && rg.is_unpack_free;
        if (geo_.is_pack_free) {
            geo_.pack_kernel = K_DUMMY;
            pack_.reset();
        } else {
            rc = pack_->create_boxes(FAM_R, es_, rg.pack_boxes);
            if (rc) return rc;
        }
        if (geo_.is_unpack_free) {  // :623
            geo_.unpack_kernel = K_DUMMY;
            unpack_.reset();
        } else {
            rc = unpack_->create_boxes(FAM_R, es_, rg.unpack_boxes);
            if (rc) return rc;
        }
    }
    std::vector<int> mapping = members;  // NCCL communicator spans the world: member -> world rank
    rc = nccl_->create(backend_, ctx_.nccl, me, mapping, geo_.send_displs, geo_.send_counts, geo_.recv_displs,
                       geo_.recv_counts, es_);
    if (rc) return rc;
    if (pipelined) nccl_->set_unpack_kernel(unpack_.get());  // null when unpack-free: nothing to run
    aux_bytes_ = nccl_->aux_bytes();
    if (geo_.is_pack_free || geo_.is_unpack_free)  // :684-686
        aux_bytes_ = std::max<int64_t>(aux_bytes_, es_ * std::max(send_elems_, recv_elems_));
    launches_ = (pack_ ? 1 : 0) + (unpack_ ? (pipelined ? P : 1) : 0);
    created_ = true;
    return DTFFT_SUCCESS;
}

int ReshapeHandle::peer_bases(void* out, std::vector<void*>* bases) {
    PeerRegistry& peers = *ctx_.peers;
    const int P = (int)members_.size();
    auto it = maps_.find(out);
    if (it != maps_.end() && it->second.id == buffer_id(out)) {
        *bases = it->second.bases;
        return DTFFT_SUCCESS;
    }
    if (it != maps_.end()) {  // the address was re-allocated: everything cached for it is stale
        fused_.erase(out);
        for (auto ci = fused_chunks_.begin(); ci != fused_chunks_.end();)
            ci = ci->first.first == out ? fused_chunks_.erase(ci) : std::next(ci);
        peers.release(&it->second.opened);
        maps_.erase(it);
        ++evictions_;
    }
    if (maps_.size() >= kMaxCachedDestinations) forget_buffers();
    PeerMap m;
    std::vector<void*> mapped;
    bool ok = false;
    int rc = peers.publish(out, (size_t)(recv_elems_ * es_), &mapped, &m.opened, &ok);
    if (rc) return rc;
    if (!ok) return DTFFTB_ERROR_NOT_REGISTERED;
    m.bases.resize((size_t)P);
    for (int i = 0; i < P; ++i) m.bases[(size_t)i] = i == me_ ? out : mapped[(size_t)members_[(size_t)i]];
    m.id = buffer_id(out);
    *bases = m.bases;
    maps_.emplace(out, std::move(m));
    return DTFFT_SUCCESS;
}

int ReshapeHandle::ensure_dma_resources() {
    cudaError_t ce;
    // at most 4: with 7 copy streams at 8 GPUs the bench's parity block caught a mismatching first backward call
    // (profiles/r02h_bench_n8_pairdma_streams7.json, unexplained); 1, 2 and 4 streams are parity- and stress-tested
    if (const char* e = getenv("DTFFTB_DMA_STREAMS")) n_copy_streams_ = std::max(1, std::min(4, atoi(e)));
    for (int i = 0; i < n_copy_streams_; ++i) {
        if (!copy_streams_[i]) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            ce = cudaStreamCreateWithPriority(&copy_streams_[i], cudaStreamNonBlocking, hi);
            if (ce != cudaSuccess) return cuda_error(ce);
            ce = cudaEventCreateWithFlags(&copies_done_[i], cudaEventDisableTiming);
            if (ce != cudaSuccess) return cuda_error(ce);
        }
    }
    size_t n_packs = 0;
    for (auto& v : dma_subs_) n_packs += v.size();
    while (pack_done_.size() < n_packs) {
        cudaEvent_t e;
        ce = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        if (ce != cudaSuccess) return cuda_error(ce);
        pack_done_.push_back(e);
    }
    return DTFFT_SUCCESS;
}

int ReshapeHandle::dma_begin(void* out, cudaStream_t stream) {
    if (!dma_mode()) return DTFFTB_ERROR_INTERNAL;
    int rc = ensure_dma_resources();
    if (rc) return rc;
    std::vector<void*> bases;
    rc = peer_bases(out, &bases);  // collective on the first use of `out`; an identity check afterwards
    if (rc) return rc;
    dma_bases_ = &maps_.find(out)->second.bases;
    for (int c = 0; c < kCopyStreams; ++c) copy_used_[c] = false;
    return ctx_.peers->barrier(members_, 2 * (comm_id_ - 1), stream);  // every member's `out` is free
}

int ReshapeHandle::dma_send(const void* in, void* out, void* aux, int peer, int sub, cudaStream_t stream) {
    (void)out;
    if (!dma_mode() || !dma_bases_ || peer < 0 || peer >= (int)members_.size() || peer == me_) return DTFFTB_ERROR_INTERNAL;
    if (sub < 0 || sub >= (int)dma_subs_[(size_t)peer].size()) return DTFFTB_ERROR_INTERNAL;
    if (!aux) return DTFFT_ERROR_INVALID_AUX;
    const DmaSub& ds = dma_subs_[(size_t)peer][(size_t)sub];
    const DmaBlock& d = ds.blk;
    if (ds.pack_index < 0 || d.run <= 0) return DTFFT_SUCCESS;
    int rc = dma_pack_->execute(in, aux, stream, ds.pack_index + 1, false);  // slice -> staging, in destination row order
    if (rc) return rc;
    const size_t ev = (size_t)ds.pack_index;
    cudaError_t ce = cudaEventRecord(pack_done_[ev], stream);
    if (ce != cudaSuccess) return cuda_error(ce);
    // the slices of one peer travel in order on ONE copy stream (its pairwise flag counts them); peers alternate
    const int c = copy_stream_of(peer);
    cudaStream_t cs = copy_streams_[c];
    copy_used_[c] = true;
    ce = cudaStreamWaitEvent(cs, pack_done_[ev], 0);
    if (ce != cudaSuccess) return cuda_error(ce);
    cudaMemcpy3DParms p{};
    const size_t row_bytes = (size_t)(d.run * es_);
    p.srcPtr = make_cudaPitchedPtr(static_cast<char*>(aux) + (size_t)(d.pack.out_off * es_), row_bytes, row_bytes, (size_t)d.rows);
    p.dstPtr = make_cudaPitchedPtr(static_cast<char*>((*dma_bases_)[(size_t)peer]) + (size_t)(d.dst_off * es_),
                                   (size_t)(d.dst_pitch * es_), row_bytes, (size_t)d.dst_plane_rows);
    p.extent = make_cudaExtent(row_bytes, (size_t)d.rows, (size_t)d.planes);
    p.kind = cudaMemcpyDefault;
    ce = cudaMemcpy3DAsync(&p, cs);  // one strided copy deposits every row at its final address in the peer's array
    return ce == cudaSuccess ? DTFFT_SUCCESS : cuda_error(ce);
}

int ReshapeHandle::dma_self(const void* in, void* out, cudaStream_t stream) {
    if (!dma_mode()) return DTFFTB_ERROR_INTERNAL;
    return dma_self_->execute_all(in, out, stream);
}

// Tell `peer` that slice `sub` of my block has landed in its array: enqueued on the copy stream that carried it.
int ReshapeHandle::dma_signal(int peer, int sub) {
    if (!dma_mode() || peer < 0 || peer >= (int)members_.size() || peer == me_) return DTFFTB_ERROR_INTERNAL;
    return ctx_.peers->signal(members_, 6 + (comm_id_ - 1), peer, sub, copy_streams_[copy_stream_of(peer)]);
}

int ReshapeHandle::dma_wait(int source, int sub, cudaStream_t stream) {
    if (!dma_mode() || source < 0 || source >= (int)members_.size() || source == me_) return DTFFTB_ERROR_INTERNAL;
    return ctx_.peers->wait(members_, 6 + (comm_id_ - 1), source, sub, stream);
}

int ReshapeHandle::dma_end(cudaStream_t stream, bool landed_barrier) {
    if (!dma_mode()) return DTFFTB_ERROR_INTERNAL;
    for (int c = 0; c < kCopyStreams; ++c) {
        if (!copy_used_[c]) continue;
        cudaError_t ce = cudaEventRecord(copies_done_[c], copy_streams_[c]);
        if (ce != cudaSuccess) return cuda_error(ce);
        ce = cudaStreamWaitEvent(stream, copies_done_[c], 0);
        if (ce != cudaSuccess) return cuda_error(ce);
    }
    dma_bases_ = nullptr;
    if (!landed_barrier) return DTFFT_SUCCESS;
    return ctx_.peers->barrier(members_, 2 * (comm_id_ - 1) + 1, stream);  // every block has landed
}

int ReshapeHandle::execute_dma(void* in, void* out, cudaStream_t stream, void* aux) {
    const int P = (int)members_.size();
    int rc = dma_begin(out, stream);
    if (rc) return rc;
    for (int k = 1; k < P; ++k) {  // rotated order: at any time every member receives from one peer
        const int p = (me_ + k) % P;
        for (int q = 0; q < n_subs_to(p); ++q) {
            rc = dma_send(in, out, aux, p, q, stream);
            if (rc) return rc;
        }
    }
    rc = dma_self(in, out, stream);  // local HBM work beside the copies
    if (rc) return rc;
    return dma_end(stream, true);
}

int ReshapeHandle::local_piece(const void* in, void* out, int side, const Pencil& x_src, const Pencil& x_dst, int peer,
                               int sub, int nsub, cudaStream_t stream) {
    if (!is_local_transpose() || (side != 0 && side != 1) || sub < 0 || sub >= nsub) return DTFFTB_ERROR_INTERNAL;
    const std::pair<int, int> key(peer, sub);
    auto it = peer_pieces_[side].find(key);
    if (it == peer_pieces_[side].end()) {
        std::unique_ptr<Kernel> k(new Kernel);
        const std::vector<Box> one = {local_box_for_block(send_, recv_by_member_[(size_t)me_], x_src, x_dst, sub, nsub)};
        int rc = k->create_boxes(FAM_T, es_, one);
        if (rc) return rc;
        it = peer_pieces_[side].emplace(key, std::move(k)).first;
    }
    return it->second->execute_all(in, out, stream);
}

int ReshapeHandle::execute_fused(void* in, void* out, cudaStream_t stream) {
    PeerRegistry& peers = *ctx_.peers;
    std::vector<void*> bases;
    int rc = peer_bases(out, &bases);  // collective on the first use of `out`; an identity check afterwards
    if (rc) return rc;
    auto it = fused_.find(out);
    if (it == fused_.end()) {
        std::unique_ptr<Kernel> k(new Kernel);
        rc = k->create_boxes(fused_family_, es_, fused_boxes_);
        if (rc) return rc;
        k->set_abort_flag(peers.abort_flag());
        rc = k->set_peer_out(bases.data(), nullptr);
        if (rc) return rc;
        it = fused_.emplace(out, std::move(k)).first;
    }
    // channel: one pair per 1-D communicator id
    const int ch_free = 2 * (comm_id_ - 1), ch_landed = ch_free + 1;
    rc = peers.barrier(members_, ch_free, stream);  // every member's `out` is free
    if (rc) return rc;
    rc = it->second->execute_all(in, out, stream);
    if (rc) return rc;
    return peers.barrier(members_, ch_landed, stream);  // every block has landed
}

int ReshapeHandle::fused_begin(void* out, cudaStream_t stream) {
    if (!can_chunk()) return DTFFTB_ERROR_INTERNAL;
    std::vector<void*> bases;
    int rc = peer_bases(out, &bases);
    if (rc) return rc;
    return ctx_.peers->barrier(members_, 2 * (comm_id_ - 1), stream);
}

int ReshapeHandle::fused_end(cudaStream_t stream) {
    return ctx_.peers->barrier(members_, 2 * (comm_id_ - 1) + 1, stream);
}

int ReshapeHandle::fused_chunk(void* in, void* out, int k, int nchunks, int max_ctas, cudaStream_t stream) {
    if (!can_chunk() || k < 0 || k >= nchunks) return DTFFTB_ERROR_INTERNAL;
    auto key = std::make_pair((const void*)out, nchunks);
    auto it = fused_chunks_.find(key);
    if (it == fused_chunks_.end()) {
        std::vector<void*> bases;
        int rc = peer_bases(out, &bases);
        if (rc) return rc;
        std::vector<std::unique_ptr<Kernel>> ks((size_t)nchunks);
        for (int c = 0; c < nchunks; ++c) {
            const std::vector<Box> boxes = chunk_boxes(send_, recv_by_member_, c, nchunks, nullptr);
            ks[(size_t)c].reset(new Kernel);
            rc = ks[(size_t)c]->create_boxes(fused_family_, es_, boxes);
            if (rc) return rc;
            ks[(size_t)c]->set_abort_flag(ctx_.peers->abort_flag());
            rc = ks[(size_t)c]->set_peer_out(bases.data(), nullptr);
            if (rc) return rc;
        }
        it = fused_chunks_.emplace(key, std::move(ks)).first;
    }
    Kernel& kern = *it->second[(size_t)k];
    kern.set_grid_limit(max_ctas);
    long long chunk_offset = 0;  // first element of chunk k in the source pencil
    chunk_boxes(send_, recv_by_member_, k, nchunks, &chunk_offset);
    return kern.execute_all(static_cast<char*>(in) + (size_t)chunk_offset * (size_t)es_, out, stream);
}

int ReshapeHandle::execute(void* in, void* out, cudaStream_t stream, void* aux) {
    if (!created_) return DTFFT_ERROR_PLAN_NOT_CREATED;
    if (!has_exchange_) return pack_->execute(in, out, stream, 0, false);
    if (backend_ == BACKEND_NVLINK_FUSED) return dma_standalone_ ? execute_dma(in, out, stream, aux) : execute_fused(in, out, stream);
    int rc;
    if (nccl_->is_pipelined()) {
        if (!aux) return DTFFT_ERROR_INVALID_AUX;
        if (geo_.is_pack_free)  // :711-716  in -> aux exchange, aux -> out unpack
            return nccl_->execute(in, out, stream, aux);
        rc = pack_->execute_all(in, aux, stream);  // in -> aux   pack
        if (rc) return rc;
        if (geo_.is_unpack_free)  // :717-722  aux -> out exchange (received straight into `out`)
            return nccl_->execute(aux, in, stream, out);
        return nccl_->execute(aux, out, stream, in);  // :723-730  aux -> in exchange, in -> out unpack
    }
    if (geo_.is_pack_free) {  // :734-740  in -> aux exchange, aux -> out unpack
        if (!aux) return DTFFT_ERROR_INVALID_AUX;
        rc = nccl_->execute(in, aux, stream, aux);
        if (rc) return rc;
        return unpack_->execute_all(aux, out, stream);
    }
    if (geo_.is_unpack_free) {  // :742-746  in -> aux pack, aux -> out exchange
        if (!aux) return DTFFT_ERROR_INVALID_AUX;
        rc = pack_->execute_all(in, aux, stream);
        if (rc) return rc;
        return nccl_->execute(aux, out, stream, aux);
    }
    rc = pack_->execute_all(in, out, stream);  // :752  in -> out  pack
    if (rc) return rc;
    rc = nccl_->execute(out, in, stream, aux);  // :755  out -> in  exchange
    if (rc) return rc;
    return unpack_->execute_all(in, out, stream);  // :758  in -> out  unpack
}

}  // namespace dtfftb
