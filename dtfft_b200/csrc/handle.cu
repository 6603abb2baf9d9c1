// ReshapeHandle (see handle.h for the reference citations).
#include "handle.h"

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "errors.h"

namespace dtfftb {

void ReshapeHandle::destroy() {
    local_pieces_[0].clear();
    local_pieces_[1].clear();
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
        geo_ = HandleGeometry{};
        geo_.ttype = ttype, geo_.rtype = rtype, geo_.ndims = ndims;
        geo_.comm_size = P, geo_.comm_rank = me, geo_.members = members;
        geo_.has_exchange = true, geo_.is_fused = true;
        launches_ = 3;  // barrier + fused kernel + barrier
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

int ReshapeHandle::local_produce(const void* in, void* out, int k, int nchunks, cudaStream_t stream) {
    if (!is_local_transpose() || k < 0 || k >= nchunks) return DTFFTB_ERROR_INTERNAL;
    auto it = local_pieces_[0].find(nchunks);
    if (it == local_pieces_[0].end()) {
        std::vector<std::unique_ptr<Kernel>> ks((size_t)nchunks);
        for (int c = 0; c < nchunks; ++c) {
            ks[(size_t)c].reset(new Kernel);
            const std::vector<Box> one = {local_producer_box(send_, recv_by_member_[(size_t)me_], c, nchunks)};
            int rc = ks[(size_t)c]->create_boxes(FAM_T, es_, one);
            if (rc) return rc;
        }
        it = local_pieces_[0].emplace(nchunks, std::move(ks)).first;
    }
    return it->second[(size_t)k]->execute_all(in, out, stream);
}

int ReshapeHandle::local_consume(const void* in, void* out, int k, int nchunks, const std::vector<Pencil>& senders_src,
                                 cudaStream_t stream) {
    if (!is_local_transpose() || k < 0 || k >= nchunks) return DTFFTB_ERROR_INTERNAL;
    auto it = local_pieces_[1].find(nchunks);
    if (it == local_pieces_[1].end()) {
        std::vector<std::unique_ptr<Kernel>> ks((size_t)nchunks);
        for (int c = 0; c < nchunks; ++c) {
            ks[(size_t)c].reset(new Kernel);
            int rc = ks[(size_t)c]->create_boxes(FAM_T, es_, local_consumer_boxes(send_, recv_by_member_[(size_t)me_],
                                                                                   senders_src, c, nchunks));
            if (rc) return rc;
        }
        it = local_pieces_[1].emplace(nchunks, std::move(ks)).first;
    }
    return it->second[(size_t)k]->execute_all(in, out, stream);
}

int ReshapeHandle::execute(void* in, void* out, cudaStream_t stream, void* aux) {
    if (!created_) return DTFFT_ERROR_PLAN_NOT_CREATED;
    if (!has_exchange_) return pack_->execute(in, out, stream, 0, false);
    if (backend_ == BACKEND_NVLINK_FUSED) return execute_fused(in, out, stream);
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
