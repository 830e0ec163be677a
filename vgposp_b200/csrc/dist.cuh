// Internal view of the distributed-inverse handle (dist.cu) for the sharded lazy-column greedy (lazy.cu).
#pragma once
#include "dense.cuh"

constexpr int64_t DIST_TAIL_RB = 512;               // rows per partial-sum block in the tail (== lazy.cu's RB)

struct vgp_dist {
    int device = 0;
    int64_t n_pad = 0;
    double *matrix = nullptr;                   // [n_pad + tail_rows][n_pad]: the replica, then the peer-visible tail
    int64_t tail_rows = 0;
    unsigned long long *flags = nullptr;        // u64[DIST_MAX] barrier words + int error + u64 option fingerprint
    unsigned long long fingerprint = 0;         // of the options that change numerics / buffer sizes, at create
    double *peer_matrix[vgp::DIST_MAX] = {nullptr};
    int ipc = 0, connected = 0;
    cudaStream_t push_stream = nullptr;         // vgp_dist_upload_rows: peer copies of a chunk under the next upload
    cudaEvent_t push_event = nullptr;
    vgp::DistContext ctx;
    vgp::DenseWorkspace ws;
};
