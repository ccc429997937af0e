/* libptk - C ABI of the B200-native lidar-odometry step for ptudes-lab.
 *
 * This is the drop-in boundary for ONE hot path of bexcite/ptudes-lab: the per-scan
 * KISS-ICP-style registration step that KissICPWrapper drives
 * (reference: src/ptudes/kiss.py:54-131, caller src/ptudes/cli/ekf_bench.py:550-555).
 * In the reference that path crosses from Python into the third-party kiss-icp 0.2.x
 * pybind module (`kiss_icp.pybind.kiss_icp_pybind`, imported through kiss.py:7-10); the
 * entry points below are what a binding for this path binds instead.  Every entry names
 * the reference call site / kiss-icp binding it replaces.
 *
 * Conventions
 *  - every call returns 0 on success, a negative PTK_E_* code otherwise; no C++ exception
 *    crosses this boundary; ptk_last_error() gives the text of the last failure.
 *  - all floating point is float64; poses are row-major 4x4 (16 doubles); point clouds are
 *    row-major (N,3) float64 (the layout kiss.py hands to kiss-icp), timestamps (N,) float64.
 *  - `const double*` inputs/outputs that carry point data may be HOST or DEVICE pointers
 *    (the library asks cudaPointerGetAttributes); small outputs (poses, counts, stats) are
 *    HOST pointers and are valid when the call returns (the call synchronises `stream`).
 *  - `stream` is a cudaStream_t passed as void* (0 = default stream), e.g. PyTorch's
 *    torch.cuda.current_stream().cuda_stream.
 *  - a context is bound to one device and is not thread safe; distinct contexts may be
 *    driven from distinct threads.  A context owns `batch` independent sequences ("lanes");
 *    the *_batch call advances all lanes in one set of kernel launches (fleet replay).
 *  - there is no CPU fallback: without a CUDA device ptk_ctx_create fails with PTK_E_CUDA.
 */
#ifndef PTK_H_
#define PTK_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PTK_VERSION 100

#define PTK_OK 0
#define PTK_E_ARG (-1)        /* bad argument */
#define PTK_E_CUDA (-2)       /* CUDA runtime error (no device, launch failure, ...) */
#define PTK_E_CAPACITY (-3)   /* scan larger than cfg.max_points / map pool or table full */
#define PTK_E_KEYRANGE (-4)   /* voxel coordinate outside +-2^20 */
#define PTK_E_NUMERIC (-5)    /* singular normal equations / non-finite pose */
#define PTK_E_STATE (-6)      /* call not valid in the current state */

#define PTK_MAX_POINTS_PER_VOXEL 20

typedef struct ptk_ctx ptk_ctx;

/* kiss_icp.config.load_config(None, deskew=True, max_range=R) + config.data.min_range = m
 * (reference: kiss.py:40-43) plus capacities of the device-side structures. */
typedef struct ptk_config {
    double max_range;            /* data.max_range, default 100 */
    double min_range;            /* data.min_range, default 5 */
    double voxel_size;           /* mapping.voxel_size; <= 0 -> max_range / 100 */
    int max_points_per_voxel;    /* mapping.max_points_per_voxel, 1..20, default 20 */
    int deskew;                  /* data.deskew, default 1 (kiss.py:41) */
    double initial_threshold;    /* adaptive_threshold.initial_threshold, default 2.0 */
    double min_motion_th;        /* adaptive_threshold.min_motion_th, default 0.1 */
    int max_iterations;          /* ICP MAX_NUM_ITERATIONS_, default 500 */
    double convergence_eps;      /* ICP ESTIMATION_THRESHOLD_, default 1e-4 */
    int max_points;              /* capacity: points per scan, default 262144 */
    int map_capacity;            /* capacity: voxels in the local map, default 262144 */
    int batch;                   /* number of independent sequences (lanes), default 1 */
    int trace_iterations;        /* >0: keep per-iteration correspondences of that many ICP
                                    iterations for ptk_get_trace (parity tap), default 0 */
} ptk_config;

/* per-scan observables (the lists kiss.py:50-52,116-124 appends to, plus counters) */
typedef struct ptk_stats {
    int status;          /* 0 ok, 1 = zero correspondences (B.5), 2 = singular solve */
    int n_in;            /* points handed in */
    int n_range;         /* after the range filter */
    int n_ds;            /* frame_downsample size (grid 0.5 v) */
    int n_src;           /* source size (grid 1.5 v) */
    int n_voxels;        /* voxels in the local map after the update */
    int iterations;      /* ICP iterations run */
    int n_corr;          /* correspondences of the last iteration */
    double dx_norm;      /* |dx| of the last iteration */
    double sigma;        /* adaptive threshold used (kiss.py:99,124) */
    double err_dt;       /* |t| of inv(guess) @ pose (kiss.py:118) */
    double err_drot;     /* |rotvec| of inv(guess) @ pose (kiss.py:119-120) */
    int map_points;      /* points in the local map after the update */
    int icp_searches;    /* full 27-voxel searches run (the rest of n_src * iterations hit the
                            correspondence cache) */
} ptk_stats;

void ptk_default_config(ptk_config* cfg);
int ptk_version(void);

/* KissICP(config) construction (kiss.py:45). */
int ptk_ctx_create(ptk_ctx** out, int device, const ptk_config* cfg);
int ptk_ctx_destroy(ptk_ctx* ctx);
/* Forget poses, threshold state and the local map of one lane (lane < 0: all lanes). */
int ptk_reset(ptk_ctx* ctx, int lane);
const char* ptk_last_error(const ptk_ctx* ctx);

/* ---- the step: KissICPWrapper._kiss_register_frame (kiss.py:83-131) ----------------
 * deskew (kiss.py:90) -> preprocess (:93) -> voxelize (:96) -> adaptive threshold (:99) ->
 * initial guess (:102-105; `initial_guess` NULL = constant velocity model) ->
 * register_frame (:108-114) -> err metrics (:116-124) -> update_model_deviation (:128) ->
 * local_map.update (:129) -> poses.append (:130).  out_pose = the new pose. */
int ptk_register_frame(ptk_ctx* ctx, int lane, const double* xyz, const double* timestamps,
                       int n, const double* initial_guess /* 16 or NULL */,
                       double* out_pose /* 16 */, ptk_stats* stats /* nullable */, void* stream);

/* Same step for every lane of the context in one set of launches (fleet replay).
 * xyz[l], timestamps[l], n[l] per lane; guesses = batch*16 doubles or NULL;
 * has_guess[l] != 0 selects guesses[l]; out_poses = batch*16; stats = batch entries. */
int ptk_register_frame_batch(ptk_ctx* ctx, const double* const* xyz,
                             const double* const* timestamps, const int* n,
                             const double* guesses, const unsigned char* has_guess,
                             double* out_poses, ptk_stats* stats, void* stream);

/* ---- the same step fed with the sensor's RANGE image: KissICPWrapper.register_frame (kiss.py:54-74)
 * including its host-side prologue - `sel = scan.field(RANGE) != 0` (:59), `xyz = xyz_lut(scan)[sel]`
 * (:60), `timestamps = self._timestamps[sel]` (:61) - on the device: 4 bytes per pixel cross the bus
 * instead of 32.  ptk_set_sensor replaces `client.XYZLut(metadata, use_extrinsics)` (kiss.py:28-29)
 * and the per-column timestamp table (kiss.py:34-35): xyz = direction * (range * range_unit)
 * [+ offset]; `offset` and `col_timestamps` may be NULL (no offset; w / W).  Pass the SDK's
 * pre-scaled direction with range_unit = 1.  `out_index` of ptk_get_points then refers to pixel
 * indices h * W + w (ascending, i.e. the same order as the masked cloud). */
int ptk_set_sensor(ptk_ctx* ctx, int H, int W, const double* direction /* H*W*3 */,
                   const double* offset /* H*W*3 or NULL */, const double* col_timestamps /* W or NULL */,
                   double range_unit);
int ptk_register_scan(ptk_ctx* ctx, int lane, const unsigned int* range_mm /* H*W, host or device */,
                      const double* initial_guess /* 16 or NULL */, double* out_pose /* 16 */,
                      ptk_stats* stats /* nullable */, void* stream);
int ptk_register_scan_batch(ptk_ctx* ctx, const unsigned int* const* range_mm, const double* guesses,
                            const unsigned char* has_guess, double* out_poses, ptk_stats* stats, void* stream);
/* Announce the host range images of the step AFTER the next one: the next ptk_register_scan[_batch]
 * call copies them to the device on a side stream while its own kernels run, and the call after it,
 * given the same host pointers, only waits for that copy.  batch entries, NULL = skip the lane.
 * Host buffers should be pinned and must stay valid until they have been consumed. */
int ptk_prefetch_scan_batch(ptk_ctx* ctx, const unsigned int* const* range_mm);

/* ---- fleet replay over several contexts (BASELINE.json configs[4]: independent sequences) ----------------------
 * The lanes of one context advance in lock step; a fleet does not have to.  ptk_fleet_replay runs n_ctx contexts,
 * each on its own host thread (and the stream streams[g], non-default), through n_scans scans:
 *   range_mm[g][s * batch_g + l]   RANGE image of scan s, lane l of context g (host pinned or device pointers)
 *   out_poses[g][(s * batch_g + l) * 16], stats[g][s * batch_g + l]   (nullable)
 * Every step of every context is an ordinary ptk_register_scan_batch (host images are prefetched one scan ahead);
 * the contexts just are not synchronised with each other, so the ICP loop of one overlaps the streaming kernels of
 * the others.  ptk_set_icp_blocks_per_lane caps the blocks of the cooperative ICP launch per lane (0 = as many as
 * the device holds) so that the ICP launches of several contexts fit the device side by side. */
int ptk_set_icp_blocks_per_lane(ptk_ctx* ctx, int blocks);
int ptk_fleet_replay(ptk_ctx* const* ctxs, int n_ctx, const unsigned int* const* const* range_mm, int n_scans,
                     double* const* out_poses, ptk_stats* const* stats, void* const* streams);

/* ---- one sequence, voxel map sharded by hash key over several GPUs (one process + context per GPU).
 * The registration loop of kiss.py:108-114 is driven by the host between its two collectives per
 * iteration (ptudes_lab_b200/sharded.py): every rank preprocesses the same scan, searches its own
 * shard of the map, the per-point records are all-gathered, each rank builds the normal-equation
 * partial of its slice of the source, the [17][32] partial table is all-reduced (NCCL), and every
 * rank solves.  Sums are combined in the canonical tree order, so poses equal the single-GPU ones
 * bit for bit.  All buffers are DEVICE pointers of float64:
 *   records  [5][n_src]         d2, order id, target x, y, z of the nearest LOCAL map point
 *   gathered [nranks][5][n_src] the records of every rank (all-gather)
 *   partials [17][32]           column r = slice root of rank r, zero elsewhere (all-reduce SUM) */
int ptk_shard_config(ptk_ctx* ctx, int rank, int nranks);
/* The same mode WITHOUT the host in the loop (the product path on NVLink-connected GPUs): after ptk_shard_config
 * every rank exports its exchange buffer (a CUDA IPC handle, or the pointer itself for contexts of one process),
 * attaches every other rank's, and from then on the ordinary step (ptk_register_scan / ptk_register_frame, lane 0,
 * the SAME scan on every rank) does everything: the ICP kernel of each rank searches its own shard and the blocks
 * with the same index exchange their per-point (d2, order id, target) records through peer memory with stamp flags,
 * take the lexicographic minimum over the ranks and build identical normal equations - no collective, no host round
 * trip per iteration, poses bit-identical to one GPU.  Streams must be non-default (the ranks' kernels spin on each
 * other).  Replaces nothing in the reference (kiss-icp is single-process); requested by BASELINE.json north_star. */
int ptk_shard_peer_export(ptk_ctx* ctx, unsigned char* handle64 /* nullable */, void** local_ptr /* nullable */,
                          unsigned long long* bytes /* nullable */);
int ptk_shard_peer_attach(ptk_ctx* ctx, int rank, const unsigned char* handle64 /* or NULL */, void* direct_ptr /* or NULL */);
int ptk_shard_begin(ptk_ctx* ctx, int lane, const double* xyz, const double* timestamps, int n,
                    const unsigned int* range_mm /* alternative input, NULL to use xyz */,
                    const double* initial_guess, int* n_src, int* n_vox_local, void* stream);
int ptk_shard_search(ptk_ctx* ctx, int lane, int iteration, double* records, void* stream);
int ptk_shard_system(ptk_ctx* ctx, int lane, const double* gathered, int iteration, double* partials, void* stream);
int ptk_shard_solve(ptk_ctx* ctx, int lane, const double* partials, int iteration, int map_empty, int* done,
                    void* stream);
int ptk_shard_end(ptk_ctx* ctx, int lane, double* out_pose, ptk_stats* stats, void* stream);

/* ---- state the wrapper exposes (kiss.py:133-166, cli/ekf_bench.py:545-547) --------- */
int ptk_num_poses(const ptk_ctx* ctx, int lane);
int ptk_get_pose(const ptk_ctx* ctx, int lane, int index /* <0 from the end */, double* out16);
/* KissICP.get_prediction_model(): inv(poses[-2]) @ poses[-1], identity if < 2 poses. */
int ptk_get_prediction_model(const ptk_ctx* ctx, int lane, double* out16);
/* KissICP.get_adaptive_threshold() WITHOUT side effects (value the next step would use is
 * only known inside the step because ComputeThreshold mutates state; this returns the last
 * sigma used). */
double ptk_last_sigma(const ptk_ctx* ctx, int lane);
/* The three state changes of the step for callers that run it PIECEWISE, i.e. keep the body of
 * KissICPWrapper._kiss_register_frame (kiss.py:83-131) and only swap the kiss-icp objects:
 * KissICP.get_adaptive_threshold() (kiss.py:99: initial threshold until has_moved(), then
 * AdaptiveThreshold.compute_threshold(), which accumulates the last model deviation - a call with side effects, as
 * upstream), adaptive_threshold.update_model_deviation(T) (kiss.py:128) and poses.append(pose) (kiss.py:130). */
int ptk_get_adaptive_threshold(ptk_ctx* ctx, int lane, double* sigma);
int ptk_update_model_deviation(ptk_ctx* ctx, int lane, const double* T16);
int ptk_append_pose(ptk_ctx* ctx, int lane, const double* T16);

/* ---- pieces, one per kiss-icp binding the wrapper or tests reach --------------------
 * kiss_icp_pybind._deskew_scan(frame, timestamps, start_pose, finish_pose) (kiss.py:76-78,90) */
int ptk_deskew_scan(ptk_ctx* ctx, const double* xyz, const double* timestamps, int n,
                    const double* start_pose, const double* finish_pose, double* out_xyz,
                    void* stream);
/* kiss_icp_pybind._preprocess(frame, max_range, min_range) (kiss.py:93) */
int ptk_preprocess(ptk_ctx* ctx, const double* xyz, int n, double max_range, double min_range,
                   double* out_xyz, int* n_out, void* stream);
/* kiss_icp_pybind._voxel_down_sample(frame, voxel_size) (kiss.py:96); out_index (nullable)
 * receives the input index of every kept point, ascending. */
int ptk_voxel_down_sample(ptk_ctx* ctx, const double* xyz, int n, double voxel_size,
                          double* out_xyz, int* out_index, int* n_out, void* stream);

/* kiss_icp_pybind._VoxelHashMap of one lane (kiss.py:129,161; registration at :108-114) */
int ptk_map_clear(ptk_ctx* ctx, int lane, void* stream);
int ptk_map_empty(ptk_ctx* ctx, int lane);                       /* 1 empty, 0 not, <0 error */
int ptk_map_update(ptk_ctx* ctx, int lane, const double* xyz, int n, const double* pose,
                   void* stream);                                 /* _update(points, pose) */
int ptk_map_add_points(ptk_ctx* ctx, int lane, const double* xyz, int n, void* stream);
int ptk_map_remove_far(ptk_ctx* ctx, int lane, const double* origin3, void* stream);
int ptk_map_num_points(ptk_ctx* ctx, int lane, int* n_points, int* n_voxels);
int ptk_map_point_cloud(ptk_ctx* ctx, int lane, double* out_xyz, int capacity, int* n_out,
                        void* stream);                            /* _point_cloud() */
/* voxel dump for parity: keys (V,3) int32, counts (V) int32, points (V,20,3) float64 */
int ptk_map_dump(ptk_ctx* ctx, int lane, int* keys, int* counts, double* points, int capacity,
                 int* n_voxels, void* stream);
/* _get_correspondences(points, max_dist): per query the order id (offset*20+slot) of the
 * nearest map point or -1, and its coordinates. */
int ptk_map_get_correspondences(ptk_ctx* ctx, int lane, const double* xyz, int n,
                                double max_dist, int* out_order, double* out_target,
                                int* n_corr, void* stream);
/* kiss_icp_pybind._register_point_cloud(points, voxel_map, initial_guess,
 * max_correspondance_distance, kernel) (kiss.py:108-114) */
int ptk_register_point_cloud(ptk_ctx* ctx, int lane, const double* xyz, int n,
                             const double* initial_guess, double max_correspondance_distance,
                             double kernel, double* out_pose, ptk_stats* stats, void* stream);

/* ---- taps on the last step (parity / `return frame, source` of kiss.py:131) -------- */
/* which: 0 = frame_downsample, 1 = source (sensor frame, before the initial guess) */
int ptk_get_points(ptk_ctx* ctx, int lane, int which, double* out_xyz, int* out_index,
                   int capacity, int* n_out, void* stream);
/* preprocessed frame (deskewed + range filtered, input order) of the last step */
int ptk_get_frame(ptk_ctx* ctx, int lane, double* out_xyz, int capacity, int* n_out, void* stream);
/* per-iteration correspondence order ids of the last registration: out (iters, n_src) */
int ptk_get_trace(ptk_ctx* ctx, int lane, int* out_order, int capacity_iters, int* n_iters,
                  int* n_src, void* stream);

/* ---- measurement taps (bench.py roofline leg) ---------------------------------------
 * With profiling on, every kernel launch of the step is bracketed by CUDA events on the
 * caller's stream; ptk_get_profile returns, per kernel slot, the accumulated device
 * milliseconds and the number of launches since the last ptk_set_profiling call.
 * ptk_launch_count: kernels launched by this context since creation (always counted). */
#define PTK_PROF_SLOTS 12
int ptk_set_profiling(ptk_ctx* ctx, int on);
int ptk_get_profile(ptk_ctx* ctx, double* ms /* PTK_PROF_SLOTS */, long long* launches /* PTK_PROF_SLOTS */);
const char* ptk_kernel_name(int slot);   /* NULL past the last used slot */
/* clock64 cycles block 0 of the last step's ICP kernel spent per phase: cache pass, searches,
 * sums, barrier wait, tree reduction, solve */
int ptk_get_icp_phases(const ptk_ctx* ctx, int lane, long long* cycles6);
long long ptk_launch_count(const ptk_ctx* ctx);
/* bytes of control data one step moves per lane besides the scan itself: the parameter record
 * uploaded before the launches and the result record (pose, counters) read back after them */
void ptk_control_bytes(int* h2d_per_lane, int* d2h_per_lane);

/* ---- consumer of the poses: the 18-state error-state EKF of `ptudes ekf-bench`, host-native -------
 * ESEKF of src/ptudes/ins/es_ekf.py:57-329: ptk_ekf_create = ESEKF(init_grav=, init_bacc=, init_bgyr=)
 * (NULL = the reference defaults), ptk_ekf_process_imu = processImu (:191-257; the sample's dt is
 * ts - previous ts, the first sample only sets the clock), ptk_ekf_process_pose = processPose
 * (:259-329; meas_cov 6x6 row-major or NULL for the default 2 cm / 0.01 rad), ptk_ekf_get_nav / _pose =
 * .nav / .nav.pose_mat(), ptk_ekf_ts = .ts.  Plain host code (no GPU needed); one filter per sequence. */
typedef struct ptk_ekf ptk_ekf;
int ptk_ekf_create(ptk_ekf** out, const double* init_grav3, const double* init_bacc3, const double* init_bgyr3);
int ptk_ekf_destroy(ptk_ekf* f);
int ptk_ekf_process_imu(ptk_ekf* f, const double* lacc3, const double* avel3, double ts);
int ptk_ekf_process_imu_batch(ptk_ekf* f, const double* lacc /* n,3 */, const double* avel /* n,3 */,
                              const double* ts /* n */, int n);
int ptk_ekf_process_pose(ptk_ekf* f, const double* pose16, const double* meas_cov36 /* nullable */);
int ptk_ekf_get_nav(const ptk_ekf* f, double* pos3, double* att9, double* vel3, double* bias_gyr3,
                    double* bias_acc3, double* grav3);
int ptk_ekf_get_pose(const ptk_ekf* f, double* pose16);
int ptk_ekf_get_cov(const ptk_ekf* f, double* cov324);
double ptk_ekf_ts(const ptk_ekf* f);

/* ---- ingest: raw Ouster UDP packets -> LidarScan fields -------------------------------
 * The step before the path for recorded data: src/ptudes/data.py:31-77 (`OusterLidarData.withScanIdx`:
 * `_client.PacketFormat.from_info`, `_client.ScanBatcher(w, pf)`, `batch(packet, ls_write)`), fed by
 * `pcap.Pcap` / `OusterRawBagSource` (src/ptudes/utils.py:171-187, src/ptudes/bag.py:21-97).  In the
 * reference the batching is ouster-sdk's C++ ScanBatcher writing HOST LidarScan fields column by column;
 * here the raw packets of a frame cross the bus once and ONE kernel scatters them into the staggered
 * (H, W) field images in HBM, where ptk_register_scan reads the RANGE image without a round trip to
 * the host.  Packet layouts restate the published Ouster sensor UDP formats [UPSTREAM-UNVERIFIED: the
 * SDK is absent; the packet sizes 24896 / 24832 / 8448 / 33024 B for 128 beams are the known answers]. */
#define PTK_PROFILE_LEGACY 1                       /* UDPProfileLidar.PROFILE_LIDAR_LEGACY */
#define PTK_PROFILE_RNG19_RFL8_SIG16_NIR16_DUAL 2  /* dual returns */
#define PTK_PROFILE_RNG19_RFL8_SIG16_NIR16 3       /* single return */
#define PTK_PROFILE_RNG15_RFL8_NIR8 4              /* low data rate */
#define PTK_IMU_PACKET_SIZE 48

/* _client.PacketFormat.from_info(metadata): byte layout of one lidar packet */
typedef struct ptk_packet_format {
    int profile, pixels_per_column, columns_per_packet, columns_per_frame;
    int packet_header_size, col_header_size, channel_data_size, col_footer_size, packet_footer_size;
    int col_size, lidar_packet_size, packets_per_frame;
} ptk_packet_format;
int ptk_packet_format_init(ptk_packet_format* pf, int profile, int pixels_per_column, int columns_per_packet,
                           int columns_per_frame);
/* frame id of a packet (packet header, or the first column header of a LEGACY packet) */
int ptk_packet_frame_id(const ptk_packet_format* pf, const unsigned char* packet);

/* Device pointers of the fields of `n_frames` frames, frame f at element offset f * (H*W) (images) or
 * f * W (per-column headers).  range/timestamp/status/measurement_id are required, the others nullable. */
typedef struct ptk_scan_fields {
    unsigned int* range;               /* ChanField.RANGE, millimetres, staggered (H, W) */
    unsigned int* range2;              /* ChanField.RANGE2 (dual profile) */
    unsigned short* reflectivity;      /* ChanField.REFLECTIVITY */
    unsigned short* signal;            /* ChanField.SIGNAL */
    unsigned short* near_ir;           /* ChanField.NEAR_IR */
    unsigned long long* timestamp;     /* LidarScan.timestamp (W) ns */
    unsigned int* status;              /* LidarScan.status (W), bit 0 = column valid */
    unsigned short* measurement_id;    /* LidarScan.measurement_id (W) */
} ptk_scan_fields;

/* ScanBatcher for whole frames: `packets` holds n_frames * pf->packets_per_frame slots of
 * pf->lidar_packet_size bytes (HOST memory - copied to the device inside the call - or DEVICE memory);
 * a slot of zero bytes is a lost packet.  Every column with its valid bit set lands in column
 * measurement_id of its frame; columns no packet covered read 0 in every field (ScanBatcher's zero fill). */
int ptk_decode_packets(const ptk_packet_format* pf, int device, const unsigned char* packets, int n_frames,
                       const ptk_scan_fields* out, void* stream);

/* ScanBatcher(w, pf) as an object: push packets in arrival order; a packet of a newer frame closes the
 * current one (data.py:58 `if batch(packet, ls_write)`), packets of older frames are dropped, a repeated
 * packet replaces its earlier copy.  Closed frames wait in pinned host memory (`frames` of them at most)
 * until ptk_batcher_decode moves the oldest to the device and decodes it. */
typedef struct ptk_batcher ptk_batcher;
int ptk_batcher_create(ptk_batcher** out, int device /* -1: host-side grouping only */, const ptk_packet_format* pf,
                       int frames);
int ptk_batcher_destroy(ptk_batcher* b);
int ptk_batcher_push(ptk_batcher* b, const unsigned char* packet, int* frames_ready);
int ptk_batcher_flush(ptk_batcher* b, int* frames_ready);     /* end of the stream: close the partial frame (data.py:52-56) */
int ptk_batcher_decode(ptk_batcher* b, const ptk_scan_fields* out, int* frame_id, int* n_packets, void* stream);
/* the oldest closed frame's packet slots as they sit in pinned host memory (tests, host-side consumers) */
int ptk_batcher_peek(ptk_batcher* b, const unsigned char** packets, int* frame_id, int* n_packets);
int ptk_batcher_pop(ptk_batcher* b);

/* pcap.Pcap(file, meta) as far as data.py needs it: UDP payloads of a pcap or pcapng file in capture order
 * (Ethernet / VLAN / Linux cooked / raw IP link types, IPv4 with reassembly of fragmented datagrams -
 * a 128-beam lidar packet is 17 Ethernet frames). */
typedef struct ptk_pcap ptk_pcap;
int ptk_pcap_open(ptk_pcap** out, const char* path);
int ptk_pcap_close(ptk_pcap* p);
/* next UDP datagram: payload copied to buf (at most cap bytes), returns 1, 0 at the end of the file */
int ptk_pcap_next(ptk_pcap* p, unsigned char* buf, int cap, int* len, int* dst_port, double* ts);
/* LZ4 frame decompression (magic 0x184D2204; dependent or independent blocks, stored blocks, checksums skipped):
 * what `rosbag record --lz4` chunks are compressed with.  Returns 0 and the size written, PTK_E_CAPACITY if `cap`
 * is too small, PTK_E_ARG for a malformed frame. */
int ptk_lz4_frame_decompress(const unsigned char* src, unsigned long long n, unsigned char* dst, unsigned long long cap,
                             unsigned long long* out_len);
/* text of the last failure of an ingest call on this thread */
const char* ptk_ingest_last_error(void);

/* pinned host memory for callers that want fast H2D of scans */
int ptk_host_alloc(void** out, unsigned long long bytes);
int ptk_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* PTK_H_ */
