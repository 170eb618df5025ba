/*
 * plvi.h -- C ABI of libplvi_cuda.so: the B200 (sm_100a) implementation of the
 * PL-VI-ORBSLAM3 visual front-end hot path.
 *
 * The reference has no FFI/plugin seam; its seam is the C++ class interface
 * (SURVEY.md section 8(b)).  Each entry point below names the reference interface it
 * stands behind (paths relative to the reference tree).  The C++ shim in
 * pl_vi_orbslam3_b200/shim/ re-exposes the reference signatures on top of this ABI;
 * INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions: plain pointers and sizes only; every function returns an int status
 * (PLVI_OK = 0, < 0 = error) unless stated; nothing throws across the ABI; the caller
 * allocates outputs at the stated capacity and the callee reports counts.  A handle
 * owns one CUDA stream and all device scratch; it is NOT re-entrant (like the
 * reference objects, which mutate mvImagePyramid), but distinct handles may run
 * concurrently.  There is no CPU fallback: without a CUDA device every call fails
 * with PLVI_ERR_CUDA.
 */
#ifndef PLVI_H_
#define PLVI_H_

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLVI_OK 0
#define PLVI_ERR_INVALID (-1)  /* bad argument */
#define PLVI_ERR_CUDA (-2)     /* CUDA runtime error; see plvi_last_error() */
#define PLVI_ERR_CAPACITY (-3) /* batch / image larger than the handle was created for */
#define PLVI_ERR_EMPTY (-4)    /* empty image: the reference operator() returns -1 */

#define PLVI_MAX_LEVELS 12

/* cv::KeyPoint as laid out by OpenCV (28 bytes) -- filled exactly as
 * ORBextractor::operator() fills it (src/ORBextractor.cc:862-872,1131-1144). */
typedef struct plvi_keypoint {
  float x, y;     /* pt, in level-0 pixels (level coords * mvScaleFactor[octave]) */
  float size;     /* (int)(31 * mvScaleFactor[octave]) */
  float angle;    /* IC_Angle, degrees [0,360) */
  float response; /* FAST score */
  int32_t octave;
  int32_t class_id; /* -1 */
} plvi_keypoint;

/* cv::line_descriptor::KeyLine (68 bytes),
 * Thirdparty/line_descriptor/include/line_descriptor/descriptor_custom.hpp:107-146 */
typedef struct plvi_keyline {
  float angle;
  int32_t class_id;
  int32_t octave;
  float pt_x, pt_y;
  float response;
  float size;
  float startPointX, startPointY, endPointX, endPointY;
  float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
  float lineLength;
  int32_t numOfPixels;
} plvi_keyline;

const char* plvi_last_error(void);
/* number of CUDA devices visible, or PLVI_ERR_CUDA */
int plvi_device_count(void);

/* ------------------------------------------------------------------ ORB ---- */
typedef struct plvi_orb plvi_orb;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (include/ORBextractor.h:50-51, src/ORBextractor.cc:408-468).  max_width/max_height/
 * max_batch size the device scratch; device = CUDA ordinal; stream = an existing
 * cudaStream_t to run on, or NULL to let the handle create its own. */
int plvi_orb_create(plvi_orb** out, int nfeatures, float scale_factor, int nlevels,
                    int ini_th_fast, int min_th_fast, int max_width, int max_height,
                    int max_batch, int device, void* stream);
void plvi_orb_destroy(plvi_orb* h);

/* Rows to allocate per frame in kps/desc: sum over levels of max(quota+3, 8)
 * (DistributeOctTree may exceed a level's quota by up to 3, src/ORBextractor.cc:667-736). */
int plvi_orb_capacity(const plvi_orb* h);
/* getters of include/ORBextractor.h:62-82 */
int plvi_orb_levels(const plvi_orb* h);
float plvi_orb_scale_factor(const plvi_orb* h);
int plvi_orb_scale_factors(const plvi_orb* h, float* scale, float* inv_scale, float* sigma2,
                           float* inv_sigma2);
int plvi_orb_features_per_level(const plvi_orb* h, int* quota);
/* level sizes for a w x h input (ComputePyramid, src/ORBextractor.cc:1156-1157) */
int plvi_orb_level_sizes(const plvi_orb* h, int w, int h_, int* lw, int* lh);
void* plvi_orb_stream(const plvi_orb* h);

/* int ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea)
 * (include/ORBextractor.h:58-60, src/ORBextractor.cc:1068-1150) over a batch of n
 * equally sized CV_8UC1 frames in HOST memory: frame i starts at imgs + i*frame_stride,
 * rows are `stride` bytes apart.  Outputs (host): kps[n][cap], desc[n][cap][32],
 * counts[n] = keypoints per frame, mono_idx[n] = the reference operator()'s return
 * value (monoIndex).  Blocking.  The mask argument of the reference is ignored there
 * and absent here. */
int plvi_orb_extract_batch(plvi_orb* h, const uint8_t* imgs, int n, int w, int h_, int stride,
                           size_t frame_stride, int lap0, int lap1, plvi_keypoint* kps,
                           uint8_t* desc, int* counts, int* mono_idx);

/* Same, asynchronous on the handle's stream: returns once the work is enqueued; host
 * buffers must stay valid (and should be pinned) until plvi_orb_sync() returns. */
int plvi_orb_extract_batch_async(plvi_orb* h, const uint8_t* imgs, int n, int w, int h_,
                                 int stride, size_t frame_stride, int lap0, int lap1,
                                 plvi_keypoint* kps, uint8_t* desc, int* counts,
                                 int* mono_idx);
int plvi_orb_sync(plvi_orb* h);
/* The host-buffer calls copy their results back on a device-to-host stream of the handle (two device result sets: the
 * copy of call i overlaps the kernels of call i + 1).  plvi_*_results_event: the cudaEvent_t that completes when the
 * results of the LAST call are in the caller's buffers (NULL before the first call); plvi_*_sync waits for everything. */
void* plvi_orb_results_event(plvi_orb* h);
int plvi_event_synchronize(void* cuda_event);                  /* host wait */
int plvi_stream_wait_event(void* stream, void* cuda_event);    /* device-side wait of a cudaStream_t */
/* Device copies of the results of the last plvi_orb_extract_batch[_async] call ([n][capacity] keypoints and
 * descriptors, [n] counts / mono indices): valid until the next call on the handle, in stream order. */
int plvi_orb_device_results(plvi_orb* h, plvi_keypoint** d_kps, uint8_t** d_desc, int** d_counts, int** d_mono_idx);

/* Same computation with every buffer already resident in DEVICE memory (inputs and
 * outputs are device pointers); enqueued on the handle's stream, not synchronised. */
int plvi_orb_extract_batch_device(plvi_orb* h, const uint8_t* d_imgs, int n, int w, int h_,
                                  int stride, size_t frame_stride, int lap0, int lap1,
                                  plvi_keypoint* d_kps, uint8_t* d_desc, int* d_counts,
                                  int* d_mono_idx);

/* Read-back of the last batch's internals (host pointers; blocking):
 *  - pyramid level (ORBextractor::mvImagePyramid[level], include/ORBextractor.h:84),
 *    dense w_l x h_l u8 into out; blurred != 0 selects the GaussianBlur'ed working
 *    copy used for descriptors (src/ORBextractor.cc:1114-1115);
 *  - FAST grid candidates of one level (vToDistributeKeys, src/ORBextractor.cc:766-855)
 *    as packed u32 (x | y << 12 | score << 24, relative to minBorder), unordered. */
int plvi_orb_read_level(plvi_orb* h, int frame, int level, int blurred, uint8_t* out);
int plvi_orb_read_candidates(plvi_orb* h, int frame, int level, uint32_t* out, int cap,
                             int* count);
/* number of kernel launches enqueued by the last extract call */
int plvi_orb_last_launches(const plvi_orb* h);
/* The per-batch launch sequence is captured once per (batch size, geometry, buffer set) into a CUDA graph and
 * replayed with a single graph launch afterwards (off while profiling or with PLVI_GRAPHS=0).  Returns the number
 * of graph replays so far; *captures (may be NULL) receives the number of captured graphs. */
int plvi_orb_graph_stats(const plvi_orb* h, int* captures);
/* void Frame::ComputeStereoMatches() (include/Frame.h, src/Frame.cc:1228-1406) for the last batch of two extractors
 * (mpORBextractorLeft / mpORBextractorRight): reads their device-resident pyramids (mvImagePyramid), so nothing is
 * copied.  Per left keypoint: candidates = right keypoints whose row band [floor(y - 2 s), ceil(y + 2 s)] contains the
 * left row, levels +-1, uR in [uL - bf / b, uL]; best Hamming distance < (TH_HIGH + TH_LOW) / 2; 11x11 SAD over +-5 px
 * on the left keypoint's level, parabola fit; mvDepth = bf / disparity; finally matches with SAD >= 1.5 * 1.4 * median
 * are dropped.  d_u_right / d_depth: [n][stride] floats (-1 = no stereo match), d_nstereo[n] = matches kept.  Device
 * pointers; runs on the left handle's stream after the right handle's pending work. */
int plvi_orb_stereo_matches(plvi_orb* left, plvi_orb* right, int n, const plvi_keypoint* d_kps_l, const uint8_t* d_desc_l,
                            const int* d_counts_l, const plvi_keypoint* d_kps_r, const uint8_t* d_desc_r, const int* d_counts_r,
                            int stride, float mb, float mbf, float* d_u_right, float* d_depth, int* d_nstereo);
/* Host-pointer form of plvi_orb_stereo_matches for one stereo frame: mvKeys / mDescriptors / mvKeysRight /
 * mDescriptorsRight in, mvuRight / mvDepth (n_l floats each) out; synchronous.  Both handles must have just
 * extracted the two images of the frame (their pyramids are read on the device). */
int plvi_orb_stereo_matches_host(plvi_orb* left, plvi_orb* right, const plvi_keypoint* kps_l, const uint8_t* desc_l, int n_l,
                                 const plvi_keypoint* kps_r, const uint8_t* desc_r, int n_r, float mb, float mbf, float* u_right,
                                 float* depth, int* nstereo);
/* Builds the image pyramid of a batch (ORBextractor::ComputePyramid, src/ORBextractor.cc:1152-1177) on the handle's
 * stream, ahead of the extraction call: the next plvi_orb_extract_batch_device on the SAME images (pointer, n, size,
 * strides) skips its pyramid stage.  Lets a caller place the pyramid's short chained kernels where nothing else
 * competes for the SMs (before the line pipeline starts) and the rest of the sequence beside region growing. */
int plvi_orb_pyramid_device(plvi_orb* h, const uint8_t* d_imgs, int n, int w, int h_, int stride, size_t frame_stride);
/* Makes the handle's stream wait for a cudaEvent_t (e.g. plvi_line_stage_event) before the next batch. */
int plvi_orb_wait_event(plvi_orb* h, void* cuda_event);
/* Makes the handle's stream wait (one polling thread, 30 ms time limit) until a device counter has reached `target`:
 * with plvi_line_stage_counter the ORB kernels start when the blocks of the line pipeline's region-growing kernel are
 * resident (launched first, they otherwise queue behind the ORB kernels' blocks and the latency-bound chain starts late). */
int plvi_orb_wait_counter(plvi_orb* h, const int* d_counter, int target);
/* As plvi_orb_wait_event, but the wait sits INSIDE the next batch's launch sequence, behind the image pyramid: the
 * pyramid kernels (short, chained, one level from the previous) run at once, FAST and everything after it when the
 * event has completed.  One-shot: applies to the next batch only. */
int plvi_orb_wait_event_after_pyramid(plvi_orb* h, void* cuda_event);
/* Per-kernel device time of the last batch: with profiling on, a CUDA event is recorded
 * on the handle's stream after every launch; plvi_orb_profile() synchronises and returns
 * "kernel=ms;kernel=ms;..." (valid until the next call). */
int plvi_orb_set_profile(plvi_orb* h, int on);
const char* plvi_orb_profile(plvi_orb* h);

/* ---------------------------------------------------------------- lines ---- */
typedef struct plvi_line plvi_line;

/* Lineextractor::Lineextractor(lsd_nfeatures, lsd_refine, lsd_scale, nlevels, scale, extractor)
 * (include/LineExtractor.h:55, src/LineExtractor.cc:39-43).  Implemented: extractor = 0
 * (LSD); lsd_refine 0 (LSD_REFINE_NONE: what every shipped yaml uses; the batched speculative region growing), 1
 * (LSD_REFINE_STD) and 2 (LSD_REFINE_ADV): refine / rect_improve / NFA of src/LSD/lsd.cpp:784-1134 in one serial warp
 * per frame and octave; nlevels 1 or 2; 0.28 <= lsd_scale <= 1 (1.0 = no blur / resize, Examples/Stereo-Line/
 * UMA_ueye.yaml); anything else returns PLVI_ERR_INVALID.  lsd_nfeatures = 0 keeps all lines (capacity 4096 per
 * frame). */
int plvi_line_create(plvi_line** out, int lsd_nfeatures, int lsd_refine, float lsd_scale, int nlevels,
                     float scale, int extractor, int max_width, int max_height, int max_batch,
                     int device, void* stream);
/* plvi_line_create with the capacity of the band-run schedule (frames; see plvi_line_set_band_run_max) chosen at
 * creation: its buffers (~10 MB per frame at 752x480) are only allocated for that many frames.  -1 = the default
 * (min(384, max_batch), or the environment variable PLVI_LSD_BR_MAX). */
int plvi_line_create_ex(plvi_line** out, int lsd_nfeatures, int lsd_refine, float lsd_scale, int nlevels, float scale,
                        int extractor, int max_width, int max_height, int max_batch, int device, void* stream, int band_run_max);
void plvi_line_destroy(plvi_line* h);
/* rows per frame in keylines / desc / line_eq */
int plvi_line_capacity(const plvi_line* h);
int plvi_line_levels(const plvi_line* h);
void* plvi_line_stream(const plvi_line* h);
int plvi_line_last_launches(const plvi_line* h);
int plvi_line_graph_stats(const plvi_line* h, int* captures);   /* see plvi_orb_graph_stats */
/* A cudaEvent_t (owned by the handle) that every batch records on the handle's stream once its streaming kernels
 * (pyramid, Gaussian, gradient) are done and only the latency-bound region growing is left, which uses a fraction of
 * the issue slots.  A caller running the ORB pipeline of the same frames on another stream can hold it back until then
 * (plvi_orb_wait_event): the two issue-bound phases no longer share the SMs and the ORB kernels fill the slots region
 * growing leaves idle.  The reference runs both extractors as two host threads per frame (src/Frame.cc:558-561). */
void* plvi_line_stage_event(plvi_line* h);
/* Device counter of region-growing blocks that have started in the current batch (reset before the stage event is
 * recorded) and the count at which all of the last batch's blocks that fit the GPU at once are resident.  Schedules
 * without that kernel (small batches, lsd_refine > 0) set the counter to a large value. */
const int* plvi_line_stage_counter(plvi_line* h, int* target);
/* mvScaleFactor_l / mvInvScaleFactor_l / mvLevelSigma2_l / mvInvLevelSigma2_l
 * (src/LineExtractor.cc:86-101) */
int plvi_line_scale_factors(const plvi_line* h, float* scale, float* inv_scale, float* sigma2,
                            float* inv_sigma2);
/* octave image sizes (LSDDetectorC::ComputePyramid) and LSD working sizes after lsd_scale */
int plvi_line_octave_sizes(const plvi_line* h, int w, int h_, int* ow, int* oh, int* sw, int* sh);

/* void Lineextractor::operator()(image, mask, keylines, descriptors_line, keylineFunction)
 * (include/LineExtractor.h:59-61, src/LineExtractor.cc:45-117) over n equally sized
 * CV_8UC1 host frames.  Outputs (host): keylines[n][cap], desc[n][cap][32],
 * line_eq[n][cap][3] (normalised homogeneous line sp x ep, f64), counts[n]
 * (PLVI_ERR_CAPACITY in counts[i] if frame i overflowed the internal segment table).
 * A frame with no lines leaves its descriptor rows untouched, like the reference. */
int plvi_line_extract_batch(plvi_line* h, const uint8_t* imgs, int n, int w, int h_, int stride,
                            size_t frame_stride, plvi_keyline* keylines, uint8_t* desc,
                            double* line_eq, int* counts);
int plvi_line_extract_batch_async(plvi_line* h, const uint8_t* imgs, int n, int w, int h_, int stride,
                                  size_t frame_stride, plvi_keyline* keylines, uint8_t* desc,
                                  double* line_eq, int* counts);
int plvi_line_sync(plvi_line* h);
void* plvi_line_results_event(plvi_line* h);
/* Device copies of the results of the last plvi_line_extract_batch[_async] call ([n][capacity] keylines, descriptors,
 * line equations, [n] counts): valid until the next call on the handle, in stream order on plvi_line_stream().  Lets a
 * caller that received the results in host buffers run the batched searches on the device copies. */
int plvi_line_device_results(plvi_line* h, plvi_keyline** d_keylines, uint8_t** d_desc, double** d_line_eq, int** d_counts);
/* The reference's Frame constructor hands the SAME image to both extractors (src/Frame.cc:558-561).  A caller that
 * batches frames can upload them once: after plvi_line_extract_batch_async(line, imgs, ...) this call runs the ORB
 * extraction on the device copy the line handle holds of those frames (same n / w / h), results to the host buffers as
 * plvi_orb_extract_batch_async; the line handle keeps that staging buffer alive until the ORB kernels have read it. */
int plvi_orb_extract_batch_async_from_line(plvi_orb* h, plvi_line* src, int lap0, int lap1, plvi_keypoint* kps, uint8_t* desc,
                                           int* counts, int* mono_idx);
int plvi_line_share_input(plvi_line* h, void* reader_stream, const uint8_t** d_img, int* pitch, size_t* frame_stride, int* n, int* w,
                          int* hh);
int plvi_line_share_done(plvi_line* h, void* reader_stream);
/* all buffers in device memory; enqueued on the handle's stream */
int plvi_line_extract_batch_device(plvi_line* h, const uint8_t* d_imgs, int n, int w, int h_,
                                   int stride, size_t frame_stride, plvi_keyline* d_keylines,
                                   uint8_t* d_desc, double* d_line_eq, int* d_counts);
/* Read-back of LSD internals of the last batch (parity tests).  what: 0 scaled f64 image
 * (after plvi_line_set_debug(h,1)), 1 level-line angle in degrees f32 (-1024 = NOTDEF),
 * 2 gradient magnitude f64, 3 raw segments (x1,y1,x2,y2 f32; *count = number),
 * 4 pyramid octave image u8 (gaussianPyrs[octave]), 5 LBD octave image u8 (w>>octave x h>>octave;
 * binary_descriptor_custom.cpp:351-371), 6 its Sobel gradients s16 {dx, dy} per pixel (:374-399),
 * 7 diagnostics of the small-batch (band-run) region growing: 40 ints = {serial fallback taken, fixed point reached,
 * -, bands re-run in round 1, 2, ...}; *count = rounds launched. */
int plvi_line_set_debug(plvi_line* h, int on);
/* Region growing has two exact schedules.  Band-run rounds (many bands per frame, warp per band, rounds to the fixed
 * point): lowest latency for a small batch (one 752x480 frame: 5 ms), throughput-bound near 8 k frames/s.  Band
 * speculation + serial commit: a latency floor of ~55 ms per batch whatever its size, but several batches in flight on
 * different handles overlap well (8 x 512 frames: 15.6 k frames/s, one batch of 4096: 22.8 k frames/s).  Batches of up to
 * max_frames frames take the band-run schedule (default: min(384, max_batch); 0 switches it off); returns the value
 * in effect.  A caller that keeps many mid-size batches in flight sets a small value (frontend.py uses 32). */
int plvi_line_set_band_run_max(plvi_line* h, int max_frames);
int plvi_line_set_profile(plvi_line* h, int on);
const char* plvi_line_profile(plvi_line* h);
int plvi_line_read_lsd(plvi_line* h, int frame, int octave, int what, void* out, int cap, int* count);

/* ------------------------------------------------------- Hamming searches ---- */
typedef struct plvi_matcher plvi_matcher;

/* Frame grid parameters: Frame::mnMinX, mnMinY, mfGridElementWidthInv,
 * mfGridElementHeightInv (src/Frame.cc:163-170; 64 x 48 cells, include/Frame.h:47-48). */
typedef struct plvi_grid {
  float min_x, min_y, inv_w, inv_h;
} plvi_grid;

/* Pinhole camera + distortion for cv::undistortPoints(src, dst, K, distCoeffs, noArray(), P):
 * K = (fx, fy, cx, cy), dist = (k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4, tx, ty) zero padded
 * (the tilt terms tx, ty must be 0), P = (new_fx, new_fy, new_cx, new_cy) - the reference passes
 * P = mK.  iters <= 0 selects OpenCV's default (5 iterations). */
typedef struct plvi_camera {
  double fx, fy, cx, cy;
  double dist[14];
  double new_fx, new_fy, new_cx, new_cy;
  int32_t iters;
} plvi_camera;

/* void Frame::UndistortKeyPoints() (src/Frame.cc:1124-1159): d_out[f][i] = d_in[f][i] with pt
 * undistorted; a plain copy when dist[0] == 0 (as the reference).  Device pointers, [n_frames][stride]
 * records, d_counts[f] valid ones; runs on `stream` (cudaStream_t) of the current device.
 * d_out may alias d_in. */
int plvi_undistort_keypoints(void* stream, const plvi_keypoint* d_in, const int* d_counts, int n_frames, int stride,
                             const plvi_camera* cam, plvi_keypoint* d_out);
/* void Frame::UndistortKeyLines() (src/Frame.cc:1161-1197): startPoint and endPoint undistorted;
 * the other KeyLine fields are copied (the reference leaves them default-constructed). */
int plvi_undistort_keylines(void* stream, const plvi_keyline* d_in, const int* d_counts, int n_frames, int stride,
                            const plvi_camera* cam, plvi_keyline* d_out);
/* void Frame::AssignFeaturesToGrid() + PosInGrid (src/Frame.cc:644-675,1077-1087), mono path:
 * mGrid[i][j] = d_cell_items[f][d_cell_start[f][i*48+j] .. d_cell_start[f][i*48+j+1]) in keypoint
 * order.  d_cell_start: [n_frames][64*48+1] ints, d_cell_items: [n_frames][stride] ints. */
int plvi_assign_features_to_grid(void* stream, const plvi_keypoint* d_keys, const int* d_counts, int n_frames, int stride,
                                 const plvi_grid* grid, int* d_cell_start, int* d_cell_items);

/* int LineMatcher::SerachForInitialize(Frame&, Frame&, vector<pair<int,int>>&) (src/LineMatcher.cpp:113-141,
 * factor 0.5) and int LineMatcher::SearchForTriangulation(KeyFrame*, KeyFrame*, vector<pair<size_t,size_t>>&)
 * (:143-171, factor 0.1, has_line1/has_line2 = "GetMapLine(idx) != NULL" flags, may be NULL), with the threshold
 * of Frame/KeyFrame::lineDescriptorMAD (src/Frame.cc:1089-1113).  Device pointers only, runs on the matcher's
 * stream.  matches12[pair][i] = matched row of desc2 or -1 (the reference's pair list is (i, matches12[i]) in
 * ascending i), mad[pair] = {nn_mad, nn12_mad}.  Pairs with fewer than two rows in desc2 give no match
 * (the reference's knnMatch(k=2) result would be read out of bounds). */
int plvi_line_match_mad(plvi_matcher* m, int npairs, const uint8_t* desc1, const int* n1, int stride1, const uint8_t* desc2,
                        const int* n2, int stride2, const uint8_t* has_line1, const uint8_t* has_line2, double factor,
                        int* matches12, int* nmatches, double* mad);
/* void MapPoint::ComputeDistinctiveDescriptors() (src/MapPoint.cc:330-402) for n_points map points:
 * d_desc [n_points][stride][32] observed descriptors (d_counts[p] valid), d_best_idx[p] = index of the
 * descriptor with the least median Hamming distance to the others (-1 if none), d_best_desc (optional)
 * [n_points][32] = that descriptor (mDescriptor).  Device pointers, `stream` = cudaStream_t. */
int plvi_distinctive_descriptors(void* stream, const uint8_t* d_desc, const int* d_counts, int n_points, int stride,
                                 int* d_best_idx, uint8_t* d_best_desc);

/* ------------------------------------------------------------- vocabulary ---- */
/* DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (ORBVocabulary, include/ORBVocabulary.h;
 * Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h).  Nodes in id order as loadFromTextFile (:1338-1424)
 * creates them: node 0 = root, parent[i] < n_nodes, the children of a node are the nodes naming it as
 * parent in id order, word ids are given to the leaves in id order.  desc: 32 bytes per node,
 * weight: WordValue (double) per node, scoring / weighting: DBoW2::ScoringType / WeightingType as in
 * the first line of ORBvoc.txt ("k L scoring weighting"). */
typedef struct plvi_vocab plvi_vocab;
int plvi_vocab_create(plvi_vocab** out, int k, int L, int scoring, int weighting, int n_nodes, const int* parent,
                      const uint8_t* is_leaf, const uint8_t* desc, const double* weight, int device);
void plvi_vocab_destroy(plvi_vocab* v);
int plvi_vocab_words(const plvi_vocab* v);
/* void TemplatedVocabulary::transform(features, BowVector& v, FeatureVector& fv, int levelsup)
 * (TemplatedVocabulary.h:1126-1194) as called by Frame::ComputeBoW (src/Frame.cc:1115-1122, levelsup 4)
 * for n_frames frames of d_counts[f] descriptors ([n_frames][stride][32] bytes).  Device pointers.
 *   per feature [n_frames][stride]: d_word_id, d_word_weight (0 = stopped word), d_node_id
 *   BowVector (std::map<WordId, WordValue>): d_bow_count[f] entries, ascending word id, in
 *     d_bow_words / d_bow_values [n_frames][stride], normalised as ScoringObject::mustNormalize says
 *   FeatureVector (std::map<NodeId, vector<unsigned>>): d_fv_count[f] nodes, ascending, in d_fv_nodes
 *     [n_frames][stride]; node r owns d_fv_features[f][d_fv_start[f][r] .. d_fv_start[f][r+1])
 *     (d_fv_start: [n_frames][stride + 1]), feature indices ascending. */
int plvi_bow_transform(plvi_vocab* v, void* stream, const uint8_t* d_desc, const int* d_counts, int n_frames, int stride,
                       int levelsup, int* d_word_id, double* d_word_weight, int* d_node_id, int* d_bow_count,
                       int* d_bow_words, double* d_bow_values, int* d_fv_count, int* d_fv_nodes, int* d_fv_start,
                       int* d_fv_features);

/* One projected query point of a guided search (28 bytes). */
typedef struct plvi_query {
  float u, v;        /* projection (uv) or vbPrevMatched[i1] */
  float radius;      /* th*mvScaleFactors[octave] | r*th*scale[level] | windowSize */
  int32_t min_level, max_level; /* GetFeaturesInArea level filter (-1 = open) */
  float angle;       /* keypoint angle of the query, for the rotation histogram */
  int32_t flags;     /* bit0: skip (no map point / outlier / octave>0 at init);
                        bit1: map point without observations (a match does not block) */
} plvi_query;

#define PLVI_SEARCH_FRAME 0     /* ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) src/ORBmatcher.cc:1962 */
#define PLVI_SEARCH_MAPPOINTS 1 /* ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, ...) src/ORBmatcher.cc:44 */
#define PLVI_SEARCH_INIT 2      /* ORBmatcher::SearchForInitialization src/ORBmatcher.cc:706 */
#define PLVI_SEARCH_BOW 3       /* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) src/ORBmatcher.cc:269 (plvi_search_by_bow) */

/* max_train <= 65535 keypoints per frame. stream: existing cudaStream_t or NULL. */
int plvi_matcher_create(plvi_matcher** out, int max_pairs, int max_train, int max_query, int device,
                        void* stream);
void plvi_matcher_destroy(plvi_matcher* m);
void* plvi_matcher_stream(const plvi_matcher* m);
int plvi_matcher_last_launches(const plvi_matcher* m);

/* static int ORBmatcher::DescriptorDistance(a, b) (include/ORBmatcher.h:43,
 * src/ORBmatcher.cc:2350-2366) and LineMatcher::distance (src/LineMatcher.cpp:173-189)
 * over n descriptor pairs (row i of a vs row i of b, 32 bytes each).  shift25 != 0
 * gives LineMatcher::DescriptorDistance (src/LineMatcher.cpp:487-499), which sums
 * floor(popcount/2) per 32-bit word.  on_device: pointers are device pointers (16-byte
 * aligned) and the call only enqueues; otherwise host pointers, blocking. */
int plvi_hamming256(plvi_matcher* m, const uint8_t* a, const uint8_t* b, int n, int shift25, int* out,
                    int on_device);

/* Guided (windowed) Hamming search over npairs independent frame pairs.  The geometry
 * that precedes the search in the reference (projection, frustum tests) stays with the
 * caller, who passes one plvi_query per candidate map point / keypoint in the
 * reference's iteration order; the train side is a Frame: undistorted keypoints
 * (mvKeysUn) + descriptors (+ optional "already has a map point with observations"
 * flags), from which the 64x48 grid (AssignFeaturesToGrid) is rebuilt on the device.
 *   train_keys [npairs][train_stride], train_desc [npairs][train_stride][32],
 *   train_blocked [npairs][train_stride] or NULL, train_counts [npairs];
 *   queries [npairs][query_stride] (INIT: u,v updated like vbPrevMatched),
 *   query_desc [npairs][query_stride][32], query_counts [npairs];
 *   th_dist: TH_HIGH (100) / TH_LOW (50); nnratio: mfNNratio; check_orientation:
 *   mbCheckOrientation (rotation histogram + ComputeThreeMaxima); the value 2 = checked, and a train keypoint whose
 *   assignment the histogram removed reads -2 instead of -1 in match_train (the reference sets mvpMapPoints[i] = NULL
 *   there, whatever it held before: src/ORBmatcher.cc:2166-2172);
 *   match_train [npairs][train_stride]: query index assigned to each train keypoint or
 *   -1 (CurrentFrame.mvpMapPoints / vnMatches21); match_query [npairs][query_stride]:
 *   train index per query or -1 (vnMatches12); nmatches [npairs]: the return value.
 * PLVI_SEARCH_FRAME with check_orientation = 0 is also the search of
 *   int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints,
 *       vector<MapPoint*>& vpMatched, int th, float ratioHamming) and its vpPointsKFs overload
 *       (include/ORBmatcher.h:53-57, src/ORBmatcher.cc:473-704),
 * which CLAIMS features while it iterates (vpMatched[idx] != NULL is skipped, :554-555; vpMatched[bestIdx] = pMP, :575):
 * train_blocked = vpMatched[i] != NULL on entry, one query per surviving map point (u, v, radius = th *
 * mvScaleFactors[nPredictedLevel], levels nPredictedLevel-1 .. nPredictedLevel, flags 0), th_dist = floor(TH_LOW *
 * ratioHamming).  Pinned against the reference's compiled ORBmatcher.cc (tests/test_oracle_vs_ref_matchers.py). */
int plvi_search_by_projection(plvi_matcher* m, int mode, int npairs, const plvi_keypoint* train_keys,
                              const uint8_t* train_desc, const uint8_t* train_blocked,
                              const int* train_counts, int train_stride, const plvi_grid* grid,
                              plvi_query* queries, const uint8_t* query_desc, const int* query_counts,
                              int query_stride, int th_dist, float nnratio, int check_orientation,
                              int* match_train, int* match_query, int* nmatches, int on_device);

/* Descriptor part of int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>&)
 * (include/ORBmatcher.h:56, src/ORBmatcher.cc:269-471, mono path).  The walk over the two
 * DBoW2::FeatureVectors stays with the caller, who passes (a) group_items
 * [npairs][items_stride]: the frame's feature indices grouped by vocabulary node, each
 * group in vIndicesF order, and (b) one plvi_query per keyframe feature of a common node,
 * in the reference's iteration order, with min_level/max_level = [start, end) of the node's
 * group inside group_items, angle = keyframe keypoint angle, flags bit0 = no / bad map
 * point.  TH_LOW, mfNNratio, rotation histogram as in the reference.  items_stride <=
 * max_train.  match_train = vpMapPointMatches as query indices. */
int plvi_search_by_bow(plvi_matcher* m, int npairs, const plvi_keypoint* train_keys, const uint8_t* train_desc,
                       const int* train_counts, int train_stride, const int* group_items, int items_stride,
                       const plvi_query* queries, const uint8_t* query_desc, const int* query_counts,
                       int query_stride, int th_dist, float nnratio, int check_orientation, int* match_train,
                       int* match_query, int* nmatches, int on_device);

/* Descriptor part of int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)
 * (include/ORBmatcher.h:57, src/ORBmatcher.cc:823-963, mono path).  Same calling scheme as plvi_search_by_bow
 * with pKF2 as the "train" side: group_items = pKF2's feature indices grouped by vocabulary node, one query per
 * pKF1 feature of a common node (flags bit0 = no / bad map point), train_blocked[i2] != 0 for pKF2 features
 * without a (good) map point (may be NULL).  th_low = TH_LOW: this variant accepts bestDist1 < TH_LOW.
 * match_query[q] = pKF2 feature matched to query q (vpMatches12[idx1] = vpMapPoints2[match_query]) or -1. */
int plvi_search_by_bow_kf(plvi_matcher* m, int npairs, const plvi_keypoint* train_keys, const uint8_t* train_desc,
                          const uint8_t* train_blocked, const int* train_counts, int train_stride, const int* group_items,
                          int items_stride, const plvi_query* queries, const uint8_t* query_desc, const int* query_counts,
                          int query_stride, int th_low, float nnratio, int check_orientation, int* match_train,
                          int* match_query, int* nmatches, int on_device);

/* Geometry of one keyframe pair for ORBmatcher::SearchForTriangulation (monocular pinhole path): F12 row-major
 * as Pinhole::epipolarConstrain builds it (K1^-T [t12]x R12 K2^-1, src/CameraModels/Pinhole.cpp:137-140), ep =
 * projection of pKF1's camera centre into pKF2 (src/ORBmatcher.cc:971-976), pKF2->mvScaleFactors /
 * mvLevelSigma2, coarse = bCoarse, check_epipole = 1 for the mono case (!bStereo1 && !bStereo2). */
typedef struct plvi_epipolar {
  float F12[9];
  float ep_x, ep_y;
  float scale_factors[16];
  float level_sigma2[16];
  int32_t coarse, check_epipole;
} plvi_epipolar;

/* Descriptor + epipolar part of int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12,
 * vector<pair<size_t,size_t>>& vMatchedPairs, bool bOnlyStereo = false, bool bCoarse) (include/ORBmatcher.h:64-65,
 * src/ORBmatcher.cc:965-1206).  Calling scheme of plvi_search_by_bow_kf: pKF2 is the searched side (train_blocked
 * = has a map point), one query per pKF1 feature WITHOUT map point of a common vocabulary node (u, v = kp1.pt,
 * min_level/max_level = the node's group range, angle, flags bit0 = skip).  Device pointers only, runs on the
 * matcher's stream.  match_query[q] = pKF2 feature or -1; vMatchedPairs = {(idx1(q), match_query[q])}. */
int plvi_search_for_triangulation(plvi_matcher* m, int npairs, const plvi_keypoint* train_keys, const uint8_t* train_desc,
                                  const uint8_t* train_blocked, const int* train_counts, int train_stride, const int* group_items,
                                  int items_stride, const plvi_query* queries, const uint8_t* query_desc, const int* query_counts,
                                  int query_stride, const plvi_epipolar* geometry, int th_low, int check_orientation,
                                  int* match_query, int* nmatches);

/* The per-map-point search shared by
 *   int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th, const bool bRight)
 *       (include/ORBmatcher.h:76, src/ORBmatcher.cc:1399-1610, mono: chi2 = 5.99, th_dist = TH_LOW),
 *   int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>&
 *       vpReplacePoint) (include/ORBmatcher.h:79, src/ORBmatcher.cc:1612-1734: chi2 = 0, TH_LOW),
 *   int ORBmatcher::SearchBySim3(KeyFrame*, KeyFrame*, vector<MapPoint*>&, s12, R12, t12, th)
 *       (include/ORBmatcher.h:71, src/ORBmatcher.cc:1736-1960: once per direction, chi2 = 0, TH_HIGH; the mutual
 *       agreement test of :1944-1957 compares the two best_idx arrays).
 * (SearchByProjection(KeyFrame*, Scw, ...) is NOT one of them: it claims features while iterating, see
 * plvi_search_by_projection.)
 * The caller projects every candidate map point (u, v, radius = th * mvScaleFactors[nPredictedLevel], min_level =
 * nPredictedLevel - 1, max_level = nPredictedLevel, flags bit0 = skipped by the checks before the search); the
 * kernel enumerates KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:1200-1244), applies the level test, the optional
 * mono reprojection gate e2 * inv_level_sigma2[level] > chi2 (chi2 <= 0: none) and keeps the first smallest Hamming
 * distance.  Queries are independent.  best_idx[q] = keyframe feature when bestDist <= th_dist, else -1;
 * best_dist[q] = the smallest distance seen (256: no candidate); nfound[pair] = number of accepted queries.  What
 * is done with a hit (Replace / AddObservation / AddMapPoint) is map bookkeeping and stays with the caller.
 * Device pointers only; inv_level_sigma2 is a host array of 16 floats; runs on the matcher's stream. */
int plvi_search_in_radius(plvi_matcher* m, int npairs, const plvi_keypoint* train_keys, const uint8_t* train_desc,
                          const int* train_counts, int train_stride, const plvi_grid* grid, const plvi_query* queries,
                          const uint8_t* query_desc, const int* query_counts, int query_stride, const float* inv_level_sigma2,
                          double chi2, int th_dist, int* best_idx, int* best_dist, int* nfound);

/* The per-map-line search of int LineMatcher::Fuse(KeyFrame* pKF, const vector<MapLine*>& vpMapLines, const float th)
 * (include/LineMatcher.h, src/LineMatcher.cpp:373-485): the caller projects the two endpoints of every candidate map
 * line and predicts its level; queries = 6 floats each (u1, v1, u2, v2, radius = th * mvScaleFactors[level], level),
 * query_flags[q] != 0 = skipped by the checks before the search (may be NULL).  Candidates =
 * KeyFrame::GetLinesInArea(u1, v1, u2, v2, radius) (src/KeyFrame.cc:1170-1198, tests exactly as written there), level
 * in [level - 1, level]; distance = LineMatcher::DescriptorDistance, i.e. the ">> 25" variant (src/LineMatcher.cpp:
 * 487-499); first smallest distance; best_idx[q] = keyline index when <= th_low (TH_LOW), else -1; best_dist[q] = the
 * smallest distance (INT_MAX: no candidate).  Replace / AddObservation / AddMapLine stay with the caller.  Device
 * pointers; runs on the matcher's stream. */
int plvi_line_fuse_search(plvi_matcher* m, int npairs, const plvi_keyline* keylines, const uint8_t* desc, const int* counts,
                          int stride, const float* queries, const uint8_t* query_flags, const uint8_t* query_desc,
                          const int* query_counts, int query_stride, int th_low, int* best_idx, int* best_dist, int* nfound);

/* Test / benchmark utility (device pointers only): builds the plvi_query records of
 * SearchByProjection(Frame,Frame) for an identity pose -- every keypoint of the query
 * frame projects onto its own position (u,v = pt), radius = th * scale_factor^octave,
 * levels octave-1..octave+1 (src/ORBmatcher.cc:2014-2023).  With a real pose the caller
 * fills plvi_query from its own projection code instead. */
int plvi_queries_from_keypoints(plvi_matcher* m, const plvi_keypoint* d_kps, const int* d_counts, int npairs,
                                int stride, float th, float scale_factor, plvi_query* d_queries);

/* Test / benchmark utilities for frame PAIRS stored as consecutive frames (query frame 2p, searched frame 2p + 1;
 * SURVEY.md section 8(d), input C3), device pointers only.  plvi_pair_queries builds, per pair, (a) the plvi_query records of
 * SearchByProjection(CurrentFrame, LastFrame) for a known image-to-image affine map in the place of the pose:
 * u = a0 x + a1 y + a2, v = a3 x + a4 y + a5, flags bit0 for points that land outside bounds4 = {min_x, max_x, min_y,
 * max_y} (dropped by the reference, src/ORBmatcher.cc:2007-2010), radius = th * scale_factor^octave, levels octave-1 ..
 * octave+1; (b) optionally (d_queries_init != NULL) the query set of SearchForInitialization (src/ORBmatcher.cc:717-727:
 * level-0 keypoints, window init_window around their own position); and the per-pair counts q_count[p] =
 * counts[2p], t_count[p] = counts[2p + 1].  d_kps is [2 * npairs][stride]; both query arrays are [npairs][out_stride]
 * (out_stride = 2 * stride lets the search address the frames of a pair in place).  plvi_gather_i32:
 * dst[i] = src[first + i * step] (per-pair count arrays of other per-frame counts). */
int plvi_pair_queries(plvi_matcher* m, const plvi_keypoint* d_kps, const int* d_counts, int npairs, int stride,
                      int out_stride, float th, float scale_factor, const float* affine6, const float* bounds4, float init_window,
                      plvi_query* d_queries_proj, plvi_query* d_queries_init, int* d_qcount, int* d_tcount);
int plvi_gather_i32(void* stream, const int* d_src, int n, int first, int step, int* d_dst);

/* static int LineMatcher::match(desc1, desc2, nnr, matches_12) (include/LineMatcher.h:87-107,
 * src/LineMatcher.cpp:92-111) with mutual != 0, LineMatcher::matchNNR (:41-61) with
 * mutual == 0, on a fresh matches_12, over npairs descriptor-set pairs:
 * BFMatcher(NORM_HAMMING).knnMatch(k=2), accept d0 < d1*nnr, both directions, keep
 * mutual matches.  With fewer than 2 train rows the reference reads out of bounds; here
 * such a pair yields no matches.  matches12 [npairs][stride1], nmatches [npairs]. */
int plvi_line_match(plvi_matcher* m, int npairs, const uint8_t* desc1, const int* n1, int stride1,
                    const uint8_t* desc2, const int* n2, int stride2, float nnr, int mutual,
                    int* matches12, int* nmatches, int on_device);

/* The stereo line search of Frame::ComputeStereoMatches_Lines (src/Frame.cc:1421-1448) over npairs left / right
 * line sets: GridStructure(grid_rows, grid_cols) filled with every right line along ORB_SLAM3::LineIterator
 * (src/gridStructure.cpp:14-23,49-60, src/LineIterator.cpp:9-52; the reference uses FRAME_GRID_ROWS x
 * FRAME_GRID_COLS = 48 x 64 and inv_width = 64 / cols, inv_height = 48 / rows, src/Frame.cc:208-209), then
 * static int LineMatcher::matchGrid(lines1, desc1, grid, desc2, directions2, w, matches_12)
 * (include/LineMatcher.h:101, src/LineMatcher.cpp:191-272) on a fresh matches_12: candidates = right lines
 * passing a cell of the window [x - win_left, x + win_right] x [y - win_up, y + win_down] around the left
 * line's start or end cell (the reference: 7, 0, 2, 2), |cos| of the directions >= 0.75, pre-emption through
 * distances[] / matches_21[], best < 0.9 * second best, mutual check.  seg = (startPointX, startPointY,
 * endPointX, endPointY) per keyline, [npairs][stride][4] floats; desc [npairs][stride][32]; grid at most
 * 64 x 64 cells.  Device pointers only; runs on the matcher's stream.  matches12 [npairs][stride1] (-1 = none),
 * nmatches [npairs] = the reference's return value.  The depth / disparity filter that follows in
 * ComputeStereoMatches_Lines is plvi_line_stereo_depth. */
int plvi_line_match_grid(plvi_matcher* m, int npairs, const float* d_seg1, const uint8_t* d_desc1, const int* d_n1,
                         int stride1, const float* d_seg2, const uint8_t* d_desc2, const int* d_n2, int stride2,
                         double inv_width, double inv_height, int grid_rows, int grid_cols, int win_left, int win_right,
                         int win_up, int win_down, int* d_matches12, int* d_nmatches);

/* What follows the search in Frame::ComputeStereoMatches_Lines (src/Frame.cc:1453-1500) for npairs stereo pairs, on the
 * matcher's stream with device pointers: for every matched left line the end-point disparities on the right line
 * carried to the left end points' rows (:1466-1470, incl. the order in which the reference overwrites sp_r / ep_r),
 * Frame::filterLineSegmentDisparity (:1535-1546, ratio 0.7), Frame::lineSegmentOverlapStereo (:1502-1533) and the
 * acceptance test (disparities >= 1, |dy| > 0.1, overlap > 0.75): disparity [npairs][stride1][2] = mvDisparity_l,
 * depth [npairs][stride1][2] = mvDepth_l = mbf / disparity, both (-1, -1) when rejected or unmatched; ndepth [npairs] =
 * lines with depth.  le (optional) [npairs][stride1][3] = mvle_l: the normalised image line through the UNDISTORTED end
 * points seg1_un (:1495-1500; zeros for a pair without right lines, where the reference returns early).  seg as in
 * plvi_line_match_grid; matches12 = its result. */
int plvi_line_stereo_depth(plvi_matcher* m, int npairs, const float* d_seg1, const int* d_n1, int stride1, const float* d_seg2,
                           const int* d_n2, int stride2, const int* d_matches12, const float* d_seg1_un, float mbf, float* d_disparity,
                           float* d_depth, double* d_le, int* d_ndepth);
/* the same for ONE stereo pair with host arrays */
int plvi_line_stereo_depth_host(plvi_matcher* m, const float* seg1, int n1, const float* seg2, int n2, const int* matches12,
                                const float* seg1_un, float mbf, float* disparity, float* depth, double* le, int* ndepth);

/* The same search for ONE stereo pair with HOST buffers (what Frame::ComputeStereoMatches_Lines holds,
 * src/Frame.cc:1421-1448): copies in, runs k_line_match_grid on the matcher's stream, copies matches12 [n1] and
 * *nmatches back, synchronises.  n1 == 0 or n2 == 0 yields no matches (the reference returns early,
 * src/Frame.cc:1419-1420). */
int plvi_line_match_grid_host(plvi_matcher* m, const float* seg1, const uint8_t* desc1, int n1, const float* seg2,
                              const uint8_t* desc2, int n2, double inv_width, double inv_height, int grid_rows,
                              int grid_cols, int win_left, int win_right, int win_up, int win_down, int* matches12,
                              int* nmatches);

/* ---- Host-buffer forms of the device-pointer searches (what the C++ shim's reference-signature adapters call:
 * shim/src/ORBmatcher.cc, shim/src/LineMatcher.cpp).  One frame / keyframe pair per call, plain host arrays in and out,
 * blocking; the inputs are staged in one stream-ordered device allocation.  Same semantics as the device forms. ---- */

/* Rectified-stereo side information for the NEXT plvi_search_by_projection (modes FRAME / MAPPOINTS) or
 * plvi_search_in_radius call on this matcher (it is consumed by that call): train_uright [npairs][train_stride] =
 * Frame::mvuRight / KeyFrame::mvuRight of the searched frame, query_ur [npairs][query_stride] = right-image
 * coordinate of every query's projection (uv.x - mbf * invz, MapPoint::mTrackProjXR).
 *   FRAME / MAPPOINTS: a candidate with mvuRight > 0 is skipped when |ur - mvuRight| > radius
 *     (src/ORBmatcher.cc:91-96, 2041-2047);
 *   plvi_search_in_radius with chi2 > 0 (ORBmatcher::Fuse): a candidate with mvuRight >= 0 is gated with
 *     (ex^2 + ey^2 + er^2) * invSigma2 > 7.8 instead of the monocular 5.99 test (src/ORBmatcher.cc:1530-1556).
 * NULL pointers clear the side information.  on_device = 0: host arrays (copied now). */
int plvi_matcher_set_stereo(plvi_matcher* m, const float* train_uright, const float* query_ur, int npairs, int train_stride,
                            int query_stride, int on_device);

/* plvi_search_in_radius for one keyframe with host arrays (ORBmatcher::Fuse x2, SearchBySim3; src/ORBmatcher.cc:1399-1960).
 * train_uright / query_ur: optional stereo side information (see plvi_matcher_set_stereo), NULL for monocular. */
int plvi_search_in_radius_host(plvi_matcher* m, const plvi_keypoint* train_keys, const uint8_t* train_desc, int n_train,
                               const plvi_grid* grid, const plvi_query* queries, const uint8_t* query_desc, int n_query,
                               const float* inv_level_sigma2, double chi2, int th_dist, const float* train_uright,
                               const float* query_ur, int* best_idx, int* best_dist, int* nfound);

/* plvi_search_for_triangulation for one keyframe pair with host arrays (src/ORBmatcher.cc:965-1206).  Rectified stereo:
 * query flags bit2 = bStereo1 (pKF1->mvuRight[idx1] >= 0), train_blocked bit1 = bStereo2; the epipole-distance test only
 * runs when neither is set (:1073-1081).  bOnlyStereo is applied by the caller (skip flag / blocked bit0). */
int plvi_search_for_triangulation_host(plvi_matcher* m, const plvi_keypoint* train_keys, const uint8_t* train_desc,
                                       const uint8_t* train_blocked, int n_train, const int* group_items, int n_items,
                                       const plvi_query* queries, const uint8_t* query_desc, int n_query,
                                       const plvi_epipolar* geometry, int th_low, int check_orientation, int* match_query,
                                       int* nmatches);

/* plvi_line_fuse_search for one keyframe with host arrays (LineMatcher::Fuse, src/LineMatcher.cpp:373-485). */
int plvi_line_fuse_search_host(plvi_matcher* m, const plvi_keyline* keylines, const uint8_t* desc, int n, const float* queries,
                               const uint8_t* query_flags, const uint8_t* query_desc, int n_query, int th_low, int* best_idx,
                               int* best_dist, int* nfound);

/* plvi_line_match_mad for one descriptor-set pair with host arrays (LineMatcher::SerachForInitialize, factor 0.5;
 * LineMatcher::SearchForTriangulation, factor 0.1 with has_line masks; src/LineMatcher.cpp:113-171). */
int plvi_line_match_mad_host(plvi_matcher* m, const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, const uint8_t* has_line1,
                             const uint8_t* has_line2, double factor, int* matches12, int* nmatches, double* mad);

/* plvi_distinctive_descriptors with host arrays: MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:330-402) and
 * MapLine::ComputeDistinctiveDescriptors (src/MapLine.cc:305-358; same median-of-distances rule on LBD descriptors with
 * LineMatcher::distance = the plain popcount).  desc [n_points][stride][32], counts [n_points]. */
int plvi_distinctive_descriptors_host(plvi_matcher* m, const uint8_t* desc, const int* counts, int n_points, int stride,
                                      int* best_idx, uint8_t* best_desc);

/* int LineMatcher::matchGrid(const std::vector<line_2d>& lines1, const cv::Mat& desc1, const GridStructure& grid,
 *     const cv::Mat& desc2, const std::vector<std::pair<double,double>>& directions2, const GridWindow& w,
 *     std::vector<int>& matches_12)   (include/LineMatcher.h:101, src/LineMatcher.cpp:191-272)
 * with exactly the data that signature carries: lines1 [n1][4] = the left lines' integer cells (sp.first, sp.second,
 * ep.first, ep.second); occ2 [n2][grid_rows][2] = per right line and grid row the first / last column of the cells of
 * `grid` that list the line ((255, 0) = none; the cells of a digital line in one grid row are contiguous, see
 * plvi_line_match_grid); directions2 [n2][2].  Host arrays, one stereo pair, blocking. */
int plvi_line_match_grid_occ_host(plvi_matcher* m, const int* lines1, const uint8_t* desc1, int n1, const uint8_t* occ2,
                                  const double* directions2, const uint8_t* desc2, int n2, int grid_rows, int grid_cols,
                                  int win_left, int win_right, int win_up, int win_down, int* matches12, int* nmatches);

/* static int ORBmatcher::DescriptorDistance(a, b) (src/ORBmatcher.cc:2350-2366) / LineMatcher::distance
 * (src/LineMatcher.cpp:173-189) / LineMatcher::DescriptorDistance (:487-499, shift25 != 0) for ONE pair on the host:
 * a pure function of two 32-byte rows that the reference calls 10^4-10^5 times per frame from MapPoint.cc:378,
 * MapLine.cc:305 and Frame.cc:1303 -- it must not cost a kernel launch.  plvi_hamming256 is the batched device form. */
static inline int plvi_inline_hamming256(const void* a, const void* b, int shift25) {
  const unsigned char* pa = (const unsigned char*)a;
  const unsigned char* pb = (const unsigned char*)b;
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t x, y;
    memcpy(&x, pa + 4 * i, 4);
    memcpy(&y, pb + 4 * i, 4);
    const int c = __builtin_popcount(x ^ y);
    dist += shift25 ? (c >> 1) : c;
  }
  return dist;
}

#ifdef __cplusplus
}
#endif
#endif /* PLVI_H_ */
