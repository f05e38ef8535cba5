/*
 * irp_b200.h -- C ABI of libirp_b200.so: the B200 (sm_100a) implementation of the embedding-based
 * outlier-detection stage of Eaglewing89/image-recognition-pipeline (functions/data_curation.py:654-728).
 *
 * The reference has no FFI of its own: its boundary is four Python functions.  The host-side mirror
 * (image-recognition-pipeline_b200/functions/data_curation.py) keeps those signatures and calls the entry
 * points below through ctypes, registered as torch custom ops (namespace irp_b200).  INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns an irp_status (0 = ok) and never throws; irp_last_error() gives the message of the
 *     most recent failure on the calling thread;
 *   - pointers named d_* are DEVICE pointers owned by the caller, h_* are host pointers; nothing returned
 *     aliases library-owned device memory except through the opaque handles;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is enqueued on it and the
 *     calls do not synchronise unless stated;
 *   - one handle must not be used from two threads at once (the reference is single-threaded,
 *     functions/data_curation.py:661-684).
 */
#ifndef IRP_B200_H_
#define IRP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum irp_status {
  IRP_OK = 0,
  IRP_ERR_INVALID = 1, /* bad argument / unsupported shape */
  IRP_ERR_CUDA = 2,    /* a CUDA runtime or driver call failed */
  IRP_ERR_NOMEM = 3,
  IRP_ERR_DEVICE = 4 /* not an sm_100 device */
} irp_status;

#define IRP_B200_ABI_VERSION 1

/* ABI version of the loaded library (IRP_B200_ABI_VERSION at build time). */
int irp_abi_version(void);
/* First 16 hex digits of the sha256 over the library's sources (csrc/*.cu, *.cuh, *.h and this header, in sorted
 * order) at build time: lets a caller check that the .so it loaded was built from the sources beside it. */
const char* irp_build_id(void);
/* Message of the last failure on this thread ("" if none). The pointer stays valid until the next failure. */
const char* irp_last_error(void);
/* Checks that `device` is a compute-capability 10.x GPU and resolves the driver entry points. */
int irp_init(int device);

/* ------------------------------------------------------------------------------------------------------------
 * A1  preprocessing  --  replaces `transform(img)` at functions/data_curation.py:675, i.e.
 * ResNet50_Weights.DEFAULT.transforms() (torchvision/transforms/_presets.py ImageClassification.forward):
 * resize shorter side -> 232 with Pillow's antialiased bilinear filter (8-bit fixed point, horizontal then
 * vertical pass, each rounded to uint8), center crop 224, /255, (x-mean)/std.  Bit-exact with Pillow before the
 * final bf16 rounding.
 *
 * Input: a ragged batch of decoded RGB images, HWC uint8, packed back to back in d_pixels; image i starts at
 * byte d_offsets[i] and has d_hw[2*i] rows and d_hw[2*i+1] columns.
 * Output layout:
 *   IRP_LAYOUT_NCHW    bf16 [n,3,224,224]            (what the reference transform returns, as bf16)
 *   IRP_LAYOUT_NHWC4P  bf16 [n,230,230,4]            (3-pixel zero border, 4th channel zero: the layout the
 *                                                     stem convolution's TMA view reads; border and pad
 *                                                     channel are (re)written by this call)
 * max_taps bounds the per-output filter taps: 2*ceil(max(1, max_i short_side_i/232))+1 for the batch.
 * ---------------------------------------------------------------------------------------------------------- */
enum { IRP_LAYOUT_NCHW = 0, IRP_LAYOUT_NHWC4P = 1, IRP_LAYOUT_U8_HWC = 2 };
enum { IRP_CROP = 224, IRP_RESIZE = 232, IRP_PAD_HW = 230 };

/* Host-only helper: resized size, crop offsets and the tap bound (2*ceil(max(scale,1))+1) for one h x w image. */
int irp_preprocess_geometry(int h, int w, int* out_h, int* out_w, int* top, int* left, int* taps);
size_t irp_preprocess_workspace_bytes(int n_images, int max_taps);
/* d_pixels must be 16-byte aligned, every d_offsets[i] a multiple of 16, and the buffer readable up to the next
 * 16-byte boundary after each image's last byte (source rows are staged with 16-byte copies). */
int irp_preprocess(const uint8_t* d_pixels, const int64_t* d_offsets, const int32_t* d_hw, int n_images,
                   int max_taps, void* d_workspace, size_t workspace_bytes, void* d_out, int out_layout,
                   void* stream);

/* Synchronises `stream` and returns IRP_ERR_INVALID if the preceding irp_preprocess(_ex) call on the same
 * workspace found max_taps too small for an image (such a call writes NaN instead of truncated-filter pixels). */
int irp_preprocess_status(const void* d_workspace, int n_images, int max_taps, void* stream);

/* The same kernels with the resize geometry selected by `transform`:
 *   IRP_TRANSFORM_WEIGHTS_DEFAULT  ResNet50_Weights.DEFAULT.transforms() (above; what irp_preprocess uses)
 *   IRP_TRANSFORM_VAL_256          the classifier's validation transform, functions/dataload.py:51-56:
 *                                  Resize((256, 256)) (aspect ratio NOT kept, Pillow antialiased bilinear),
 *                                  CenterCrop(224), ToTensor, Normalize(ImageNet mean/std)
 *   IRP_TRANSFORM_WDS_LANCZOS      the WebDataset stage's resize_and_crop_image, functions/data_curation.py:
 *                                  883-913: smaller side -> 224 with Pillow's LANCZOS filter
 *                                  (support 3), the other side int(side * (224 / smaller)), crop offsets by floor
 *                                  division.  Meant for IRP_LAYOUT_U8_HWC: uint8 [n,224,224,3], the bytes of the
 *                                  PIL image the reference returns (any layout / transform pair is accepted).
 *   IRP_TRANSFORM_HASH_64          the duplicate hash's resize, functions/data_curation.py:283-292:
 *                                  img.resize((64, 64)) with Pillow's default BICUBIC filter (a = -0.5, support 2),
 *                                  aspect ratio not kept, no crop.  IRP_LAYOUT_U8_HWC only; the output is
 *                                  uint8 [n,64,64,3] (irp_md5_rows then hashes the 12 288 bytes of each image).
 *                                  max_taps: 2*ceil(2*max(1, h/64, w/64))+1.
 * max_taps for VAL_256: 2*ceil(max(1, h/256, w/256))+1; for WDS_LANCZOS: 2*ceil(3*max(1, smaller/224))+1
 * (irp_preprocess_geometry_ex returns it per image). */
enum { IRP_TRANSFORM_WEIGHTS_DEFAULT = 0, IRP_TRANSFORM_VAL_256 = 1, IRP_TRANSFORM_WDS_LANCZOS = 2,
       IRP_TRANSFORM_HASH_64 = 3 };
enum { IRP_VAL_RESIZE = 256, IRP_HASH_SIZE = 64 };
int irp_preprocess_geometry_ex(int h, int w, int transform, int* out_h, int* out_w, int* top, int* left, int* taps);
int irp_preprocess_ex(const uint8_t* d_pixels, const int64_t* d_offsets, const int32_t* d_hw, int n_images,
                      int max_taps, void* d_workspace, size_t workspace_bytes, void* d_out, int out_layout,
                      int transform, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * A0/A2  ResNet-50 trunk  --  replaces initialize_model (functions/data_curation.py:654-659) and the per-image
 * `model(img_tensor)` at :677 (torchvision/models/resnet.py:108-160,266-282 minus fc).  Eval-mode BatchNorm is
 * folded into bf16 weights + fp32 bias; ReLU / residual add run in the conv epilogues.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct irp_resnet50 irp_resnet50;

enum { IRP_RESNET50_NUM_CONVS = 53, IRP_EMBED_DIM = 2048 };

/* Creates the trunk for up to max_batch images per call (activation arena + tensor maps). */
int irp_resnet50_create(irp_resnet50** out, int max_batch);
void irp_resnet50_destroy(irp_resnet50* net);
/* Shape of conv `index` in execution order (conv1, then per block conv1,conv2,conv3[,downsample]). */
int irp_resnet50_conv_shape(int index, int* cout, int* cin, int* kh, int* kw, int* stride);
/* Loads conv `index`: fp32 OIHW weight plus its BatchNorm (gamma, beta, running_mean, running_var, eps), all
 * device pointers; folds and re-lays them out on the device. */
int irp_resnet50_load_conv(irp_resnet50* net, int index, const float* d_weight_oihw, const float* d_gamma,
                           const float* d_beta, const float* d_mean, const float* d_var, float eps, void* stream);
/* d_x: IRP_LAYOUT_NHWC4P bf16 batch; d_embed: fp32 [batch,2048] pooled embeddings (the `.squeeze()` of :677). */
int irp_resnet50_embed(irp_resnet50* net, const void* d_x_nhwc4p, int batch, float* d_embed, void* stream);
/* Same as irp_resnet50_embed, and additionally copies the NHWC bf16 output of conv `capture_index` (after its
 * fused epilogue) into d_capture_bf16 (parity hook for the per-layer tests).  Index 0 (the stem) yields the tensor
 * AFTER the 3x3/2 max pool fused into the stem kernel, [batch,56,56,64] (the unpooled stem output never exists).
 * With a capture every convolution runs as its own launch, including layer1's first shortcut convolution, which
 * the product path folds into the junction kernel's accumulator (irp_conv1x1_chain_ds has its own parity test). */
int irp_resnet50_embed_capture(irp_resnet50* net, const void* d_x_nhwc4p, int batch, float* d_embed,
                               int capture_index, void* d_capture_bf16, size_t capacity_elems, void* stream);

/* One fused convolution (the building block of the trunk), exposed for parity tests:
 * x NHWC bf16 [B,H,W,Cin], w bf16 [Cout,kh,kw,Cin], bias fp32 [Cout], residual NHWC bf16 or NULL.
 * Supported: (kh,kw,stride,pad) in {(1,1,1,0),(1,1,2,0),(3,3,1,1),(3,3,2,1)}, Cin % 64 == 0, Cout % 64 == 0. */
int irp_conv2d_nhwc(const void* d_x, const void* d_w, const float* d_bias, const void* d_residual, void* d_out,
                    int B, int H, int W, int Cin, int Cout, int ksize, int stride, int relu, void* stream);

/* Two chained 1x1 convolutions (a bottleneck's conv3 + residual + ReLU, then the next bottleneck's conv1 + ReLU),
 * the fused form the trunk uses at the layer1 / layer2 block junctions; exposed for parity tests.
 *   y  [rows,N1] = relu(t2 [rows,K1] . w3[N1,K1]^T + b3 + residual [rows,N1])
 *   t1 [rows,N2] = relu(y . w1[N2,N1]^T + b1)
 * all bf16 row-major, biases fp32.  K1 % 64 == 0, N1 % 128 == 0, N1 <= 1024, N2 in {64,128,256}. */
int irp_conv1x1_chain(const void* d_t2, const void* d_w3, const float* d_b3, const void* d_residual, void* d_y,
                      const void* d_w1, const float* d_b1, void* d_t1, int64_t rows, int K1, int N1, int N2,
                      void* stream);

/* The same junction for a block whose shortcut is a stride-1 1x1 convolution of the block input x (layer1's first
 * bottleneck, torchvision/models/resnet.py:155-156 `identity = self.downsample(x)`): the shortcut is computed in
 * the conv3 accumulator, so the downsample tensor is never written to or read back from HBM.
 *   y  [rows,N1] = relu(t2 [rows,K1] . w3^T + x [rows,K2] . wds^T + bias)     wcat = [w3 | wds]  [N1, K1+K2]
 *   t1 [rows,N2] = relu(y . w1[N2,N1]^T + b1)                                bias = b3 + bds
 * K1 % 64 == 0, K2 % 64 == 0, K2 > 0, N1 % 128 == 0, N1 <= 1024, N2 in {64,128,256}. */
int irp_conv1x1_chain_ds(const void* d_t2, const void* d_x, const void* d_wcat, const float* d_bias, void* d_y,
                         const void* d_w1, const float* d_b1, void* d_t1, int64_t rows, int K1, int K2, int N1, int N2,
                         void* stream);

/* The last bottleneck's conv3 + residual + ReLU + global average pool, the fused form the trunk ends with (the
 * 2048-channel activation is never written); exposed for parity tests:
 *   out [batch,Cout] (fp32) = mean over the 49 pixels of relu(t2 [batch*49,K] . w[Cout,K]^T + bias + residual)
 * t2 / w / residual bf16 row-major, K = 512, Cout % 128 == 0 (torchvision/models/resnet.py:150-159, :278-279). */
int irp_conv1x1_pool(const void* d_t2, const void* d_w, const float* d_bias, const void* d_residual, float* d_out,
                     int batch, int K, int Cout, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * N3  duplicate-detection hash  --  replaces compute_image_hash (functions/data_curation.py:283-292; call site
 * :394-399): md5 of the RGB bytes of img.resize((64, 64)).  irp_preprocess_ex(IRP_LAYOUT_U8_HWC,
 * IRP_TRANSFORM_HASH_64) is the resize; irp_md5_rows is RFC 1321 MD5 of every row of a [n_rows, row_bytes] uint8
 * matrix (row_bytes = 12 288 for the hash): d_digest uint8 [n_rows, 16], hexdigest = the 16 bytes in order.
 * ---------------------------------------------------------------------------------------------------------- */
int irp_md5_rows(const uint8_t* d_data, int n_rows, int64_t row_bytes, uint8_t* d_digest, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * A3  PCA  --  replaces PCA(n_components).fit_transform at functions/data_curation.py:700-701 with the exact
 * covariance route (sklearn/decomposition/_pca.py:587-640 covariance_eigh, _base.py:151-159 transform,
 * utils/extmath.py:973-981 sign convention).
 *
 *   irp_cov_accumulate : adds this shard's  n, sum(x - s), sum (x-s)(x-s)^T  into fp64 accumulators
 *                        (split-bf16 tensor-core GEMM, fp32 per-chunk accumulate, fp64 combine).  Multi-GPU:
 *                        all-reduce the three accumulators (that is the stage's only collective), then
 *   irp_pca_fit        : mean, covariance, fp64 symmetric eigensolve of the top k pairs (Lanczos with full
 *                        re-orthogonalisation, or Householder tridiagonalisation for small / rank-deficient
 *                        problems; bisection + inverse iteration on the tridiagonal), sklearn's sign convention;
 *   irp_pca_transform  : Z = (X - mean) V^T  (split-bf16 tcgen05 GEMM, fp32 accumulate).
 * ---------------------------------------------------------------------------------------------------------- */
size_t irp_cov_workspace_bytes(int64_t n_rows, int dim);
int irp_cov_accumulate(const float* d_x, int64_t n_rows, int dim, const float* d_shift, double* d_count,
                       double* d_sum, double* d_scatter, void* d_workspace, size_t workspace_bytes, void* stream);

size_t irp_pca_fit_workspace_bytes(int dim, int k);
/* Outputs (device): mean fp64[dim]; components fp64[k,dim] (row-major, sklearn sign convention); eigenvalues
 * fp64[k+1]: the k largest eigenvalues of the covariance in descending order (explained_variance_) followed by
 * the total variance trace(C) (denominator of explained_variance_ratio_). Work is enqueued on `stream`; the call
 * may synchronise `stream` (convergence check of the Lanczos iteration) and must not be stream-captured. */
int irp_pca_fit(const double* d_count, const double* d_sum, const double* d_scatter, const float* d_shift,
                int dim, int k, double* d_mean, double* d_components, double* d_eigenvalues, void* d_workspace,
                size_t workspace_bytes, void* stream);
/* irp_pca_fit with the eigensolver chosen by the caller and a host-side report.  solver: AUTO = Lanczos when it
 * applies (dim >= 512, 5 <= k <= dim/8) with the Householder path as its fallback (breakdown / no convergence);
 * LANCZOS / HOUSEHOLDER force one.  lanczos_first_check > k moves the first convergence check (0 = 3k+10 steps).
 * h_info (host, may be NULL) receives {solver that produced the result, Lanczos steps, convergence checks, 0}. */
enum { IRP_PCA_SOLVER_AUTO = 0, IRP_PCA_SOLVER_LANCZOS = 1, IRP_PCA_SOLVER_HOUSEHOLDER = 2 };
int irp_pca_fit_ex(const double* d_count, const double* d_sum, const double* d_scatter, const float* d_shift,
                   int dim, int k, double* d_mean, double* d_components, double* d_eigenvalues, void* d_workspace,
                   size_t workspace_bytes, int solver, int lanczos_first_check, int32_t* h_info, void* stream);
/* Z = (X - mean) V^T as a tensor-core GEMM: y = float(double(x) - mean) and the components are split exactly
 * into three bf16 terms each and the six leading products accumulate in fp32 (k <= 128, dim % 64 == 0). */
size_t irp_pca_transform_workspace_bytes(int64_t n_rows, int dim, int k);
int irp_pca_transform(const float* d_x, int64_t n_rows, int dim, const double* d_mean, const double* d_components,
                      int k, float* d_z, void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * A4  outlier scoring  --  replaces detect_outliers (functions/data_curation.py:709-728):
 * LocalOutlierFactor(n_neighbors, contamination).fit_predict(...) == -1, per class and globally
 * (sklearn/neighbors/_lof.py:286-332,498-523).  Brute-force fp64 Euclidean k-NN restricted to rows of the same
 * group, reachability / LRD / LOF, np.percentile threshold, strict `<` flag.
 *
 *   d_group[i] in [0, n_groups): rows are scored only against rows of their own group (per-class pass);
 *   pass d_group = NULL and n_groups = 1 for the global pass.  An id outside the range fails the call with
 *   IRP_ERR_INVALID (checked on the host: the grouped calls synchronise `stream` once).  n_neighbors is clipped
 *   to group size - 1 like sklearn (_lof.py:286-293); a group with a single member has no neighbours, where
 *   sklearn raises: its row gets score -1 and is never flagged.
 *   d_scores: negative_outlier_factor_ per row (fp64); d_offsets: threshold per group (fp64 [n_groups]);
 *   d_flags: 1 where score < threshold of the row's group.
 *
 * irp_centroid_zscore is the north_star's distance-to-centroid scorer (no reference counterpart): per group
 * centroid, Euclidean distance, z-score of the distance, flag = distance above the (1-contamination) percentile.
 * ---------------------------------------------------------------------------------------------------------- */
size_t irp_lof_workspace_bytes(int64_t n_rows, int dim, int k);
int irp_lof(const float* d_z, int64_t n_rows, int dim, const int32_t* d_group, int n_groups, int k,
            double contamination, double* d_scores, double* d_offsets, uint8_t* d_flags, void* d_workspace,
            size_t workspace_bytes, void* stream);

/* Multi-GPU form of irp_lof: every rank holds all rows (after the all-gather of the projected
 * rows) and runs the O(n^2) neighbour search for 1/n_parts of the query tiles only; what the other ranks need are three
 * length-n fp64 vectors, each exchanged with ONE sum all-reduce (rows a rank does not own are written as 0):
 *
 *   irp_lof_knn_part   -> d_kdist   [n]  k-distance of the rows this part owns          (all-reduce)
 *   irp_lof_lrd_part   -> d_lrd     [n]  local reachability density of the owned rows   (all-reduce)
 *   irp_lof_score_part -> d_score   [n]  negative_outlier_factor_ of the owned rows     (all-reduce)
 *   irp_lof_finish     -> scores in input order, per-group np.percentile offset, flags  (every rank, O(n))
 *
 * All vectors are in the library's group-sorted row order; the SAME workspace (irp_lof_workspace_bytes) must be
 * passed to the four calls of one problem.  irp_lof is these four calls with n_parts = 1. */
int irp_lof_knn_part(const float* d_z, int64_t n_rows, int dim, const int32_t* d_group, int n_groups, int k, int part,
                     int n_parts, double* d_kdist, void* d_workspace, size_t workspace_bytes, void* stream);
int irp_lof_lrd_part(int64_t n_rows, int n_groups, int k, int part, int n_parts, const double* d_kdist_all,
                     double* d_lrd, void* d_workspace, size_t workspace_bytes, void* stream);
int irp_lof_score_part(int64_t n_rows, int n_groups, int k, int part, int n_parts, const double* d_lrd_all,
                       double* d_score_sorted, void* d_workspace, size_t workspace_bytes, void* stream);
int irp_lof_finish(int64_t n_rows, int n_groups, int k, double contamination, const double* d_score_sorted_all,
                   double* d_scores, double* d_offsets, uint8_t* d_flags, void* d_workspace, size_t workspace_bytes,
                   void* stream);

size_t irp_centroid_workspace_bytes(int64_t n_rows, int dim, int n_groups);
int irp_centroid_zscore(const float* d_z, int64_t n_rows, int dim, const int32_t* d_group, int n_groups,
                        double contamination, double* d_dist, double* d_zscore, double* d_thresholds,
                        uint8_t* d_flags, void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * N1  classifier inference (first "next" row)  --  the head of AnimalClassifier
 * (functions/model.py:29-40: Dropout, Linear(in_dim, hidden), ReLU, Dropout, Linear(hidden, num_classes); eval
 * mode, dropouts are identities) on the trunk's pooled features, and the statistics evaluate_full accumulates per
 * batch (functions/train.py:208-216).  Inputs come from irp_preprocess_ex(IRP_TRANSFORM_VAL_256) ->
 * irp_resnet50_embed.  All fp32, row-major, torch.nn.Linear weight layout [out, in].
 *
 *   irp_classifier_head     : logits [batch, num_classes]; pred [batch] = argmax (first maximum), may be NULL
 *   irp_cross_entropy_stats : d_stats[0] = sum_i w[y_i] * CE_i, d_stats[1] = sum_i w[y_i] (w = 1 when
 *                             d_class_weights is NULL), d_stats[2] = #(argmax_i == y_i); nn.CrossEntropyLoss's
 *                             batch mean is stats[0] / stats[1].  Labels outside [0, num_classes) are skipped.
 * ---------------------------------------------------------------------------------------------------------- */
size_t irp_classifier_head_workspace_bytes(int batch, int hidden);
int irp_classifier_head(const float* d_features, int batch, int in_dim, const float* d_w1, const float* d_b1,
                        int hidden, const float* d_w2, const float* d_b2, int num_classes, float* d_logits,
                        int32_t* d_pred, void* d_workspace, size_t workspace_bytes, void* stream);
int irp_cross_entropy_stats(const float* d_logits, const int64_t* d_labels, int batch, int num_classes,
                            const float* d_class_weights, double* d_stats, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * N4  UMAP graph construction (last "next" row)  --  the part of
 * `umap.UMAP(**umap_params).fit_transform(features_pca, y=y_numeric)` (functions/data_curation.py:704-705) that turns
 * the PCA output into UMAP's weighted k-NN graph; umap-learn 0.5.7 (requirements.txt:175; not vendored):
 *
 *   irp_knn_graph          : umap_.py nearest_neighbors, EXACT (Euclidean, fp64 distances, ties by index -- the
 *                            neighbour search of the LOF scorer).  Row i of d_idx / d_dist [n_rows, k] = i itself
 *                            (distance 0) followed by its k-1 nearest other rows, ascending; k = UMAP's n_neighbors.
 *                            umap-learn searches exactly below 4 096 samples and with NN-descent above; the arrays
 *                            are what UMAP(precomputed_knn=(idx, dist)) takes.
 *   irp_umap_fuzzy_weights : umap_.py smooth_knn_dist (rho, sigma by binary search; n_iter 64, bandwidth 1 and
 *                            local_connectivity 1 are UMAP's defaults) + compute_membership_strengths: d_vals
 *                            [n_rows, k] = weight of the directed edge i -> d_idx[i, j] (0 for i itself), float32
 *                            arithmetic like umap-learn's numba kernels.  The symmetrisation A + A^T - A.A^T, the
 *                            categorical intersection with the labels and the layout stay host-side with UMAP.
 *                            Workspace: 8 bytes.
 * parity unpinned: umap-learn cannot be installed in the build container; oracle/umap_graph_ref.py restates the
 * published functions and is pinned to sklearn's brute-force neighbours and to the defining equations only.
 * ---------------------------------------------------------------------------------------------------------- */
size_t irp_knn_graph_workspace_bytes(int64_t n_rows, int dim, int k);
int irp_knn_graph(const float* d_z, int64_t n_rows, int dim, int k, int32_t* d_idx, float* d_dist, void* d_workspace,
                  size_t workspace_bytes, void* stream);
int irp_umap_fuzzy_weights(const int32_t* d_idx, const float* d_dist, int64_t n_rows, int k, float local_connectivity,
                           float bandwidth, int n_iter, float* d_sigma, float* d_rho, float* d_vals, void* d_workspace,
                           size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IRP_B200_H_ */
