/*
 * nf_b200.h -- C ABI of libnf_b200.so: B200 (sm_100a) kernels for NeuroFluid's two hot paths.
 *
 * The reference (syguan96/NeuroFluid) has no FFI layer of its own: its hot paths hide behind two
 * Python module classes and three third-party operator call sites.  Each
 * entry point below names the reference interface it replaces (paths relative to the reference
 * repository).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C: pointers, sizes, scalars; no torch / C++ types cross the boundary.
 *   - every pointer is a DEVICE pointer into caller-owned memory unless its name ends in _host.
 *   - every call is asynchronous on the CUDA stream passed as `stream` (a cudaStream_t cast to
 *     void*; NULL = legacy default stream).  No call synchronises the device, with two exceptions:
 *     nf_render_backward waits once for two row counters of its forward (they size the backward's tile
 *     loops), and nf_comm_init blocks inside ncclCommInitRank.
 *   - no hidden device allocation: scratch comes from the caller, sized by the *_bytes() queries
 *     (nf_comm_init creates an NCCL communicator, which allocates its own buffers).
 *   - return value: 0 = ok, < 0 = error (NF_E_*); nf_last_error() returns a thread-local message.
 *   - fp32 everywhere at the boundary; neighbour indices int32; neighbour counts int64 where the
 *     reference returns int64 tensors.
 */
#ifndef NF_B200_H
#define NF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NF_API __attribute__((visibility("default")))
#else
#define NF_API
#endif

#define NF_B200_VERSION 100

#define NF_OK 0
#define NF_E_INVALID (-1)   /* bad argument (null pointer, unsupported size)            */
#define NF_E_WORKSPACE (-2) /* caller workspace too small                               */
#define NF_E_CUDA (-3)      /* a CUDA runtime call failed; see nf_last_error()          */
#define NF_E_UNSUPPORTED (-4)

/* operand dtype of the tensor-core contractions (accumulation is always fp32) */
#define NF_DTYPE_F16 0
#define NF_DTYPE_BF16 1

NF_API int nf_version(void);
NF_API const char* nf_last_error(void);
/* number of kernel launches issued by this library since process start (bench bookkeeping) */
NF_API int64_t nf_launch_count(void);
/* Optional per-stage device timing of nf_render_forward (bench bookkeeping): while enabled every call
 * records CUDA events on its stream between stages; nf_profile_read() synchronises on them, returns the
 * summed milliseconds of [ray query coarse, MLP coarse, composite+resample+ray query fine, MLP fine,
 * composite fine] and the number of calls covered, and clears the record. */
NF_API int nf_profile_enable(int on);
NF_API int nf_profile_read(double* stage_ms_host /*[5]*/, int* n_calls_host);

/* ---------------------------------------------------------------------------------------------
 * Spatial grid (shared by both neighbour searches)
 * replaces: the O(R*S*P) scan + particles.repeat(R,1,1) of models/renderer.py:113-118 and
 *           Open3D's FixedRadiusSearch hash-table build inside ContinuousConv (models/transmodel.py:116)
 * cell: edge length of a grid cell (use >= 1.002 * search radius).
 * ------------------------------------------------------------------------------------------- */
NF_API size_t nf_grid_workspace_bytes(int n_points);
NF_API int nf_grid_build(const float* pos /*(n,3)*/, int n_points, float cell, void* grid_ws, size_t grid_ws_bytes,
                  void* stream);

/* ---------------------------------------------------------------------------------------------
 * First-K-by-index ball query   (parity / debug entry point)
 * replaces: pytorch3d.ops.ball_query(p1, p2, K=, radius=) at models/renderer.py:116-118
 * idx_out (nq,K) int32: the K smallest particle indices with |q-p|^2 < radius^2, ascending,
 *   -1 padded (pytorch3d returns the same set in the same order);  count_out (nq) int32 =
 *   min(#in-radius, K).  Squared distances / gathered neighbours are recomputed by the consumer.
 * ------------------------------------------------------------------------------------------- */
NF_API int nf_ballquery_firstk(const void* grid_ws, int n_points /* as passed to nf_grid_build */,
                               const float* queries /*(nq,3)*/, int nq, float radius, int K,
                        int32_t* idx_out, int32_t* count_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * NeRF MLP weights   replaces: models/nerf.py:57-81 parameter storage (nn.Linear fp32)
 * params: 24 device pointers per net in this order (weight then bias for each):
 *   xyz_encoding_1..8, xyz_encoding_final, dir_encoding, sigma, rgb       (row-major (out,in) fp32)
 * packed_out: nf_render_packed_weights_bytes() bytes; layout is private (tcgen05 operand tiles).
 * ------------------------------------------------------------------------------------------- */
NF_API size_t nf_render_packed_weights_bytes(void);
NF_API int nf_render_pack_weights(const float* const* params_host /*[24] device pointers*/, int dtype, void* packed_out,
                           void* stream);
/* Encoding ablations (cfg.encoding.{density, smoothed_pos, var, smoothed_dir}, models/renderer.py:152-175): a network built
 * without a feature block has narrower first / skip / direction layers (in_xyz = 63 + 9 d + 63 s + 63 v, in_dir = 27 + 27 sd).
 * The kernels always produce all six encodings; the packer places the narrower weight matrices into the full column layout
 * and leaves the columns of a disabled block zero, which is the same function.  enc_flags = OR of NF_ENC_*. */
#define NF_ENC_DENSITY 1
#define NF_ENC_SMOOTHED_POS 2
#define NF_ENC_VAR 4
#define NF_ENC_SMOOTHED_DIR 8
#define NF_ENC_ALL 15
NF_API int nf_render_pack_weights_ex(const float* const* params_host, int dtype, int enc_flags, void* packed_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused positional-encoding + NeRF MLP over compact geometry records (parity / debug entry point)
 * replaces: Embedding.forward x6 (models/nerf.py:21-38, called from models/renderer.py:125-179)
 *           + NeRF.forward (models/nerf.py:83-124)
 * records (n_rows,16) fp32: [x(3), density, smoothed(3), variance(3), ray_dir(3), smoothed_dir(3)]
 * out (n_rows,4) fp32: [r,g,b,sigma]   (sigma_only != 0: [0,0,0,sigma], layers after sigma skipped)
 * workspace: nf_nerf_mlp_workspace_bytes() of device memory, 256-byte aligned (per-CTA staging of the encoded features
 *            between the warps that compute them and the bulk copies that feed them to the tensor cores; a constant ~45 MB,
 *            independent of n_rows; nf_render_forward carves the same region out of its own workspace)
 * ------------------------------------------------------------------------------------------- */
NF_API size_t nf_nerf_mlp_workspace_bytes(void);
NF_API int nf_nerf_mlp_forward(const void* packed_weights, int dtype, const float* records, int n_rows, int sigma_only,
                        float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole renderer forward over one chunk of rays
 * replaces: RenderNet.forward (models/renderer.py:211-270), .coarse_rendering (:273-307),
 *           .fine_rendering (:310-369) incl. utils/ray_utils.py coarse_sample_ray /
 *           ImportanceSampling / sample_pdf and models/renderer.py render_image.
 * ------------------------------------------------------------------------------------------- */
#define NF_RENDER_FORWARD 0 /* coarse + fine                                   */
#define NF_RENDER_COARSE 1  /* coarse only                                     */
#define NF_RENDER_FINE 2    /* sigma-only coarse pass, then fine               */

#define NF_SEARCH_AUTO 0   /* sweep when n_particles <= 65536, else stream    */
#define NF_SEARCH_STREAM 1 /* index-order stream through a cell bitmap (any P) */
#define NF_SEARCH_SWEEP 2  /* capsule gather -> index bitmap -> candidate sweep */

typedef struct nf_render_args {
    /* scene */
    const void* grid_ws;    /* nf_grid_build() of `particles` with cell >= 1.002*radius        */
    const float* particles; /* (n_particles,3)                                                 */
    int32_t n_particles;
    /* rays */
    const float* rays; /* (n_rays,6) [origin(3), direction(3)]                                  */
    int32_t n_rays;
    float ro[3]; /* camera position used for the smoothed-direction feature (set_ro)          */
    const float* ro_dev; /* optional device float[3]: read instead of ro (no host copy of a device tensor) */
    /* sampling */
    const float* z_coarse; /* (n_coarse) depths shared by all rays: near*(1-t)+far*t            */
    const float* u_importance; /* (n_importance) inverse-CDF arguments: linspace(0,1,n)         */
    int32_t n_coarse, n_importance;
    /* search */
    float radius;
    int32_t K;
    int32_t search;     /* NF_SEARCH_*: first-K search flavour                                  */
    /* behaviour */
    int32_t mode;       /* NF_RENDER_*                                                         */
    int32_t use_mask;   /* cfg.use_mask                                                        */
    int32_t white_background;
    int32_t dtype;      /* NF_DTYPE_*                                                          */
    const void* weights_coarse; /* packed                                                      */
    const void* weights_fine;
    /* outputs (any may be NULL): shapes as in the reference's result dict */
    float* rgb0;      /* (n_rays,3) */
    float* depth0;    /* (n_rays)   */
    float* opacity0;  /* (n_rays)   */
    int64_t* num_nn0; /* (n_rays,n_coarse)   */
    float* mask0;     /* (n_rays)   */
    float* rgb1;
    float* depth1;
    float* opacity1;
    int64_t* num_nn1; /* (n_rays,n_coarse+n_importance) */
    float* mask1;
    /* scratch */
    void* workspace;
    size_t workspace_bytes;
    /* optional statistics written by the device (may be NULL): int32[16] =
       {MLP rows coarse, MLP rows fine, active samples coarse, active samples fine,
        fine pass: group scans, solo row-scan queries, scan steps / 64, candidates tested / 64,
        coarse pass: the same four, 4 reserved} */
    int32_t* stats;
    int32_t flags; /* NF_RENDER_SAVE_NEIGHBORS: keep every record row's neighbour list in the workspace (parity tests,
                      backward pass); size the workspace with nf_render_workspace_bytes_ex(..., K, flags) */
    /* training-time jitter (utils/ray_utils.py:245-253, 186-190; models/renderer.py:192-196): random numbers are the caller's */
    int32_t z_stride;    /* 0: z_coarse is one table of n_coarse depths shared by all rays;  else z_coarse holds one row of
                            z_stride floats per ray (perturb > 0: stratified depths, ascending within a ray) */
    int32_t u_stride;    /* likewise for u_importance (perturb > 0: u ~ U[0,1) per ray and sample) */
    const float* noise0; /* optional (n_rays, n_coarse): added to sigma before the ReLU in the coarse compositing */
    const float* noise1; /* optional (n_rays, n_coarse + n_importance): the same for the fine pass */
    /* cfg.encoding.exclude_ray == False (models/renderer.py:100-109): smoothed position = x (1 - alpha) + weighted mean * alpha,
       alpha = 0.9 when same_smooth_factor, else 0.1 for samples with <= 20 neighbours and 0.9 above */
    int32_t include_ray;        /* 0: exclude_ray=True (every shipped config) */
    int32_t same_smooth_factor;
} nf_render_args;

#define NF_RENDER_SAVE_NEIGHBORS 1

/* Byte offsets of the regions of a render workspace that outlive nf_render_forward (tests compare them with the oracle,
 * nf_render_backward reads them).  Row r of rec/rowid/nbr (r < counters[0] coarse, counters[1] fine) is one evaluated
 * sample: rec (16 floats, layout of nf_nerf_mlp_forward), rowid = ray * S + sample, nbr = K neighbour indices (-1 padded). */
typedef struct nf_render_ws_view {
    size_t counters, act0, act1, z1, rec0, rowid0, out0, rec1, rowid1, out1, nbr0, nbr1, miss, total;
    int32_t act_stride0, act_stride1, cap0, cap1;
} nf_render_ws_view;

NF_API size_t nf_render_workspace_bytes(int n_rays, int n_coarse, int n_importance);
NF_API size_t nf_render_workspace_bytes_ex(int n_rays, int n_coarse, int n_importance, int K, int flags);
NF_API int nf_render_workspace_view(int n_rays, int n_coarse, int n_importance, int K, int flags, nf_render_ws_view* view_host);
NF_API int nf_render_forward(const nf_render_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Renderer backward (training)
 * replaces: autograd through RenderNet.forward, i.e. loss.backward() at trainer/trainer_e2e.py:277 and
 *           trainer/trainer_renderer.py:96: gradients w.r.t. both NeRF MLPs' parameters and w.r.t. the particle positions
 *           (through the K neighbour positions of every evaluated sample; sample positions are detached,
 *           utils/ray_utils.py:224).
 * Parameter gradients use one flat fp32 buffer per network of nf_render_param_count() floats: the 24 tensors of
 * nf_render_pack_weights in the same order (weight (out,in) row-major, then bias), concatenated.  All gradient outputs are
 * ACCUMULATED into (zero them first).  nf_render_backward synchronises the stream once (it reads the row counts).
 * ------------------------------------------------------------------------------------------- */
NF_API size_t nf_render_param_count(void);
NF_API size_t nf_render_packed_weights_bwd_bytes(void);
/* transposed bf16 weight slabs for the data-gradient GEMMs; params as for nf_render_pack_weights */
NF_API int nf_render_pack_weights_bwd(const float* const* params_host /*[24] device pointers*/, void* packed_out, void* stream);
NF_API int nf_render_pack_weights_bwd_ex(const float* const* params_host, int enc_flags, void* packed_out, void* stream);
/* Backward of nf_nerf_mlp_forward (parity / debug entry point, and the MLP stage of nf_render_backward):
 * dout4 (n,4): gradient w.r.t. (pre-sigmoid r, g, b, sigma), read at index rowid[row] (rowid NULL: row);
 * dfeat (n_rows,272) out: gradient w.r.t. the encoded features [xyz-like 198 | 10 pad | dir-like 54 | 10 pad];
 * dparams: flat parameter gradients (accumulated). */
NF_API size_t nf_nerf_mlp_backward_workspace_bytes(int n_rows);
NF_API int nf_nerf_mlp_backward(const void* packed_fwd, const void* packed_bwd, int dtype, const float* records,
                                const int32_t* rowid, const float* dout4, int n_rows, float* dfeat, float* dparams,
                                void* workspace, size_t workspace_bytes, void* stream);

typedef struct nf_render_bwd_args {
    const nf_render_args* fwd; /* the arguments of the forward call (flags & NF_RENDER_SAVE_NEIGHBORS); its workspace must
                                  be untouched since */
    const void* weights_coarse_bwd; /* nf_render_pack_weights_bwd */
    const void* weights_fine_bwd;
    /* upstream gradients, shapes as the forward outputs; any may be NULL (= zero) */
    const float* d_rgb0;
    const float* d_depth0;
    const float* d_opacity0;
    const float* d_rgb1;
    const float* d_depth1;
    const float* d_opacity1;
    /* outputs, accumulated */
    float* d_particles;     /* (n_particles,3) */
    float* d_params_coarse; /* nf_render_param_count() floats */
    float* d_params_fine;
    void* workspace;        /* nf_render_backward_workspace_bytes(..., rows of the two passes = counters[0], counters[1]) */
    size_t workspace_bytes;
} nf_render_bwd_args;
NF_API size_t nf_render_backward_workspace_bytes(int n_rays, int n_coarse, int n_importance, int rows_coarse, int rows_fine);
NF_API int nf_render_backward(const nf_render_bwd_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Transition model
 * replaces: open3d.ml.torch.layers.ContinuousConv.forward (models/transmodel.py:116,118,125),
 *           ml3d.ops.reduce_subarrays_sum (:135), nn.Linear x4 (:117,126) and
 *           ParticleNet.forward (:151-163)
 * params: device pointers, fp32, in this order:
 *   conv0_fluid.{kernel,bias}, conv0_obstacle.{kernel,bias}, dense0_fluid.{weight,bias},
 *   conv1.{kernel,bias}, dense1.{weight,bias}, conv2.{..}, dense2.{..}, conv3.{..}, dense3.{..}   (18)
 *   kernels are (4,4,4,cin,cout), dense weights (out,in).
 * ------------------------------------------------------------------------------------------- */
NF_API size_t nf_transition_packed_weights_bytes(void);
NF_API int nf_transition_pack_weights(const float* const* params_host /*[18] device pointers*/, int dtype,
                               void* packed_out, void* stream);
NF_API size_t nf_transition_workspace_bytes(int n_fluid, int n_box);

typedef struct nf_transition_args {
    const float* pos; /* (n_fluid,3) */
    const float* vel; /* (n_fluid,3) */
    int32_t n_fluid;
    const float* box;         /* (n_box,3) */
    const float* box_normals; /* (n_box,3) */
    int32_t n_box;
    float gravity[3];
    float dt;
    float filter_extent; /* diameter; search radius = extent/2 */
    int32_t dtype;
    const void* weights; /* packed */
    /* outputs */
    float* pos_out;  /* (n_fluid,3) */
    float* vel_out;  /* (n_fluid,3) */
    float* nnbr_out; /* (n_fluid) float: number of fluid neighbours */
    /* optional debug outputs (may be NULL) */
    float* feats0_out; /* (n_fluid,96) concatenated [obstacle, fluid, dense] */
    float* delta_out;  /* (n_fluid,3) position correction */
    void* workspace;
    size_t workspace_bytes;
    /* rank-sharded execution (n>1 GPUs): this rank computes particles [shard_begin, shard_end) of every
       layer; rows outside the shard must be supplied by the caller between layers (see
       neurofluid_b200/distributed.py).  Single GPU: 0, n_fluid. */
    int32_t shard_begin, shard_end;
    const void* box_grid_ws; /* optional: nf_grid_build(box, n_box, 1.002 * filter_extent / 2) done once by the caller
                                for a static container (the role of the reference's fixed_radius_search_hash_table
                                argument); NULL: built inside every step */
    int32_t* overflow_out; /* optional device int32[2]: += number of particles of this call whose fluid / box neighbour
                              list exceeded the 128 slots and was truncated (the reference has no cap): callers that
                              care poll it -- results are only reference-exact while it stays 0 */
    int32_t phase; /* -1: whole step on all particles;  NF_PHASE_SHARDED (-2): whole step with the particles block-sharded over
                      the ranks of nf_comm_init -- each rank computes rows [rank * per, (rank + 1) * per), per =
                      ceil(n_fluid / world), and the library all-gathers the three activation matrices and the packed
                      (pos, vel, count, delta) rows in place on `stream` (4 NCCL calls, no host work in between): every
                      rank ends up with the full outputs, bit-identical to the single-GPU step;
                      0..4: run only that phase on [shard_begin, shard_end):
                      0 integrate + grids + neighbour lists + layer 0,  1..3 conv layers,  4 position update.
                      Between phases the caller all-gathers the layer outputs (nf_transition_layer_buffer). */
} nf_transition_args;

#define NF_PHASE_SHARDED (-2)
NF_API int nf_transition_num_phases(void);
NF_API int nf_transition_step(const nf_transition_args* args, void* stream);
/* Location of layer `layer`'s activation matrix (the ReLU'd fp16/bf16 rows the next layer gathers from)
 * inside a transition workspace: *offset_bytes from the workspace base, *row_bytes per particle.
 * layer 0: 96 channels, 1 and 2: 64 channels.  Used by the sharded execution to all-gather rows.
 * layer 3..6: the fp32 pre-activation outputs of the four layers (models/transmodel.py:122-131 `ans_convs`): 96, 64, 64
 * floats per row and, for the last one, 16 floats per row of which the first 3 are the position correction * 128. */
NF_API int nf_transition_layer_buffer(int n_fluid, int n_box, int layer, size_t* offset_bytes_host,
                                      size_t* row_bytes_host);

/* ---------------------------------------------------------------------------------------------
 * Transition model backward (training)
 * replaces: autograd through ParticleNet.forward: loss.backward() at trainer/trainer_transmodel.py:197 (two-step unroll:
 *           gradients flow through pos and vel into the previous step) and trainer/trainer_e2e.py:277.  ContinuousConv
 *           gradients exist w.r.t. filters and input features only (as in Open3D); positions get theirs through
 *           pos_new + delta and vel = (pos_out - pos) / dt (models/transmodel.py:144-148).
 * d_params: one flat fp32 buffer of nf_transition_param_count() floats = the 18 tensors of nf_transition_pack_weights
 * in that order, concatenated; ACCUMULATED into.  d_pos / d_vel are overwritten.
 * ------------------------------------------------------------------------------------------- */
NF_API size_t nf_transition_param_count(void);
NF_API size_t nf_transition_packed_weights_bwd_bytes(void);
NF_API int nf_transition_pack_weights_bwd(const float* const* params_host /*[18] device pointers*/, void* packed_out, void* stream);
NF_API size_t nf_transition_backward_workspace_bytes(int n_fluid);
typedef struct nf_transition_bwd_args {
    const nf_transition_args* fwd; /* the forward call's arguments (phase -1); its workspace must be untouched since */
    const void* weights_bwd;       /* nf_transition_pack_weights_bwd */
    const float* g_pos_out;        /* (n_fluid,3) upstream gradient w.r.t. pos_out, or NULL */
    const float* g_vel_out;        /* (n_fluid,3) upstream gradient w.r.t. vel_out, or NULL */
    float* d_pos;                  /* (n_fluid,3) */
    float* d_vel;                  /* (n_fluid,3) */
    float* d_params;               /* nf_transition_param_count() floats, accumulated */
    void* workspace;
    size_t workspace_bytes;
} nf_transition_bwd_args;
NF_API int nf_transition_backward(const nf_transition_bwd_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU exchange (one process per GPU): NCCL over NVLink / NVSwitch, on the caller's stream
 * replaces: nothing -- the reference is single-GPU; this is the "position all-gather
 *           per step" of the particle-block sharded transition step.
 * Rank 0 creates an id (nf_comm_unique_id), the caller ships its NF_COMM_ID_BYTES bytes to every rank by whatever channel
 * it has (torch.distributed broadcast, a file, MPI), every rank calls nf_comm_init.  One communicator per process.
 * nf_allgather_rows: in-place all-gather of equal blocks -- rank g's bytes_per_rank bytes sit at buf + g * bytes_per_rank
 * on every rank; enqueued on `stream` right behind whatever produced the block (no host synchronisation).
 * ------------------------------------------------------------------------------------------- */
#define NF_COMM_ID_BYTES 128
NF_API int nf_comm_unique_id(void* id_out_host);
NF_API int nf_comm_init(const void* id_host, int rank, int world);
NF_API int nf_comm_finalize(void);
NF_API int nf_comm_info(int* rank_host, int* world_host);
/* Peer-memory exchange.  COLLECTIVE (every rank calls it, each with its own buffer of the same size; synchronises the device):
 * maps every rank's buffer into every other rank through CUDA IPC.  Afterwards an nf_allgather_rows on a range inside the
 * registered buffer -- and the four exchanges of the sharded nf_transition_step when its workspace is the registered buffer --
 * is one kernel that stores this rank's rows into every peer's copy over NVLink and raises an epoch flag there, plus a one-warp
 * kernel that waits for the peers' flags: no NCCL call on the data path.  The buffer must come from a plain cudaMalloc
 * allocation (PyTorch's default caching allocator qualifies); NF_E_UNSUPPORTED otherwise -- then every rank stays on NCCL.
 * Registering again replaces the previous buffer; buf = NULL unregisters.  nf_comm_exchange_timeouts: how many waits gave up
 * after ~100 s because a peer never arrived (0 in a healthy run; results are undefined otherwise). */
NF_API int nf_comm_register_buffer(void* buf, size_t bytes);
NF_API int nf_comm_exchange_timeouts(unsigned int* count_host);
NF_API int nf_allgather_rows(void* buf, size_t bytes_per_rank, void* stream);

/* ---------------------------------------------------------------------------------------------
 * One ContinuousConv (operator-level drop-in for a maintainer who keeps models/transmodel.py and swaps only the layer)
 * replaces: open3d.ml.torch.layers.ContinuousConv.forward as constructed at models/transmodel.py:79-98 (kernel_size
 *           [4,4,4], linear interpolation, ball_to_cube_volume_preserving, normalize=False, window = poly6 or none,
 *           radius_search_ignore_query_points) and called at :116, :118, :125; count_out is what
 *           ml3d.ops.reduce_subarrays_sum(ones, conv.nns.neighbors_row_splits) returns at :135-138.
 * Supported channel shapes: cin * cout <= 768 (fp32 on CUDA cores: the model's 4->32, 3->32, 64->3 layers) and
 * 64 -> 64 / 96 -> 64 (tcgen05, fp16/bf16 operands, fp32 accumulate); anything else: NF_E_UNSUPPORTED.
 * ------------------------------------------------------------------------------------------- */
NF_API size_t nf_cconv_packed_weights_bytes(int cin, int cout);
NF_API int nf_cconv_pack_weights(const float* kernel /*(4,4,4,cin,cout)*/, const float* bias /*(cout) or NULL*/, int cin,
                                 int cout, int dtype, void* packed_out, void* stream);
NF_API size_t nf_cconv_workspace_bytes(int n_in, int n_out, int cin, int cout);

typedef struct nf_cconv_args {
    const void* grid_in;   /* nf_grid_build(in_positions, n_in, cell >= 1.002 * extent / 2) */
    const float* in_feat;  /* (n_in, cin)  */
    int32_t n_in, cin;
    const float* out_pos;  /* (n_out, 3)   */
    int32_t n_out, cout;
    float extent;          /* filter diameter; search radius = extent / 2 (d^2 <= r^2) */
    int32_t use_window;    /* poly6 window clamp((1 - d^2/r^2)^3, 0, 1) (models/transmodel.py:73-77) */
    int32_t ignore_same;   /* radius_search_ignore_query_points: skip in-points whose coordinates equal the out-point's */
    int32_t dtype;         /* NF_DTYPE_*: operand type of the tensor-core shapes */
    const void* weights;   /* nf_cconv_pack_weights */
    float* out;            /* (n_out, cout) = conv + bias */
    float* count_out;      /* optional (n_out): number of neighbours of every out point (uncapped) */
    int32_t* nbr_index_out; /* optional (n_out, 128): neighbour indices in search order, -1 padded (conv.nns.neighbors_index) */
    int32_t* overflow_out; /* optional device int32[2]: [0] += out points with more than 128 neighbours (list truncated) */
    void* workspace;
    size_t workspace_bytes;
} nf_cconv_args;
NF_API int nf_cconv_forward(const nf_cconv_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Callers' side of the hot paths: camera rays and evaluation metrics on the device
 * ------------------------------------------------------------------------------------------- */
/* replaces: get_ray_directions + get_rays (utils/ray_utils.py:85-104, 107-130) for one view.
 * c2w_dev: (3,4) row-major camera-to-world matrix in device memory; rays_out: (H*W,6) [origin, unit direction],
 * pixel (row j, column i) at index j*W + i. */
NF_API int nf_generate_rays(int H, int W, float focal, const float* c2w_dev, float* rays_out, void* stream);
/* replaces: cKDTree(pred).query(gt) (utils/point_eval.py:11-14): for every query the Euclidean distance to (and
 * optionally the index of) the nearest point of the grid-sorted set (nf_grid_build with any cell size). */
NF_API int nf_nearest_distance(const void* grid_ws, int n_points, const float* queries, int n_queries, float* dist_out,
                               int32_t* idx_out /* may be NULL */, void* stream);
/* replaces: _distance (utils/point_eval.py:7-8): per-point distance of two equally ordered (n,3) sets. */
NF_API int nf_pair_distance(const float* a, const float* b, int n, float* dist_out, void* stream);
/* replaces: img2mse numerator (trainer/trainer_e2e.py:24): sum over n floats of (a-b)^2 into a device double. */
NF_API int nf_sqdiff_sum(const float* a, const float* b, long long n, double* sum_out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NF_B200_H */
