/*
 * j2k_b200.h — C ABI of the B200-native JPEG 2000 sample-domain path.
 *
 * This is the drop-in boundary for go-dicom-codec's data-parallel hot path
 * (DC level shift, RCT / ICT / Part-2 custom MCT, multi-level 5/3 and 9/7
 * lifting DWT in both directions, scalar quantization / dequantization).
 * T1 (EBCOT/MQ), T2, HTJ2K block coding and codestream parsing stay in Go.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, and names
 * the reference code it replaces as `path:line` relative to the reference
 * repository root (github.com/cocosip/go-dicom-codecs).  The cgo binding a
 * maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - return 0 on success, a negative j2k_status on failure; the message is
 *     available through j2k_last_error(ctx) (kept per context, not per thread).  Nothing ever
 *     falls back to a CPU implementation: without a usable CUDA device every
 *     compute entry point fails with J2K_ERR_CUDA.
 *   - all host buffers are caller-owned; no pointer is retained after return
 *     (cgo rule).  Buffers from j2k_acquire_buffer() are library-owned pinned
 *     memory and may be used for zero-staging transfers.
 *   - every entry point selects its CUDA device itself (goroutines migrate
 *     between OS threads across cgo calls); a context is thread-safe.
 *
 * Buffer layouts
 *   pixels : DICOM native frame, component-interleaved samples, 1 byte when
 *            bit_depth <= 8 else 2 bytes little-endian
 *            (jpeg2000/encoder.go:357-380, jpeg2000/decoder.go:857-859,938-940).
 *   coeffs : for each tile in raster order (jpeg2000/encoder.go:1966-1983), for
 *            each component, one row-major int32[th][tw] plane in Mallat layout
 *            (LL_L top-left; per resolution HL right, LH below, HH diagonal;
 *            rectangles as jpeg2000/encoder.go:2352-2389, jpeg2000/t2/geometry.go:53-92),
 *            row stride = tile width.  This is exactly the `tileData [][]int32`
 *            that buildTilePacketEncoder consumes (jpeg2000/encoder.go:2391) and
 *            the `comp.coefficients` that applyIDWT consumes
 *            (jpeg2000/t2/tile_decoder.go:886).
 */
#ifndef J2K_B200_H
#define J2K_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define J2K_B200_ABI_VERSION 1

#define J2K_MAX_COMPONENTS 4   /* jpeg2000/encoder.go:298 (1..4 components) */
#define J2K_MAX_LEVELS 10      /* API caps at 6 (encoder.go:306); the wavelet package and norm tables reach deeper (quantization.go:17-22) */
#define J2K_MAX_BANDS (3 * J2K_MAX_LEVELS + 1)
#define J2K_MAX_BINDINGS 4

typedef enum j2k_status {
    J2K_OK = 0,
    J2K_ERR_INVALID_ARG = -1, /* parameter validation failed (encoder.go:291-339) */
    J2K_ERR_SIZE = -2,        /* buffer too small ("insufficient pixel data", encoder.go:346-348) */
    J2K_ERR_CUDA = -3,        /* CUDA runtime / driver failure, or no device */
    J2K_ERR_NOMEM = -4,
    J2K_ERR_UNSUPPORTED = -5,
    J2K_ERR_TICKET = -6
} j2k_status;

/* Multi-component transform selector.
 * encode: Encoder.Encode dispatch, jpeg2000/encoder.go:196-209
 * decode: Decoder.applyInverseTransforms, jpeg2000/decoder.go:620-628 */
typedef enum j2k_mct_mode {
    J2K_MCT_NONE = 0,
    J2K_MCT_RCT = 1,          /* colorspace/rct.go:6-21 (3 components, reversible) */
    J2K_MCT_ICT = 2,          /* encode: float32 ICT encoder.go:277-288; decode: float64 ICT colorspace/ict.go:16-21 */
    J2K_MCT_CUSTOM_INT = 3,   /* encode: integer matrix, int64 accumulate, encoder.go:489-505 */
    J2K_MCT_CUSTOM_Q13 = 4,   /* encode: Q13 fixed point, encoder.go:506-523,662-665 */
    J2K_MCT_CUSTOM_FLOAT = 5, /* decode: float64 matrix + math.Round, decoder.go:696-723 */
    J2K_MCT_BINDINGS = 6      /* Part-2 binding list, encoder.go:527-660 / decoder.go:630-694 */
} j2k_mct_mode;

/* One Part-2 binding; mirrors MCTBindingParams (jpeg2000/encoder.go:111-121,
 * built by MCTBindingBuilder jpeg2000/mct_builder.go:4-29) on encode and the
 * decoder's mctBinding (jpeg2000/decoder.go:630-694) on decode. */
typedef struct j2k_mct_binding {
    int32_t n_components;                      /* len(ComponentIDs); 0 = all components in order (encoder.go:562-568) */
    int32_t component_ids[J2K_MAX_COMPONENTS];
    int32_t element_type;                      /* encode: 0 = integer matrix else Q13 (encoder.go:554-558); decode: 0 = integer (reversible) else float64 (decoder.go:636-641) */
    int32_t has_matrix;                        /* 0 = identity (encoder.go:596-608); decode: 0 = skip the matrix */
    double matrix[J2K_MAX_COMPONENTS * J2K_MAX_COMPONENTS]; /* row-major n x n */
    int32_t has_offsets;                       /* encode subtracts before (encoder.go:582-594), decode adds after (decoder.go:683-694) */
    int32_t offsets[J2K_MAX_COMPONENTS];
} j2k_mct_binding;

/* Forward (encode-side) parameters: each field mirrors something the Go path reads. */
typedef struct j2k_fwd_params {
    int32_t width, height;       /* EncodeParams.Width/Height        encoder.go:19-20 */
    int32_t components;          /* EncodeParams.Components          encoder.go:21 */
    int32_t bit_depth;           /* EncodeParams.BitDepth (1..16)    encoder.go:22 */
    int32_t is_signed;           /* EncodeParams.IsSigned            encoder.go:23 */
    int32_t tile_width;          /* 0 = whole image                  encoder.go:26,1990-1997 */
    int32_t tile_height;
    int32_t num_levels;          /* EncodeParams.NumLevels           encoder.go:30 (this ABI accepts up to J2K_MAX_LEVELS) */
    int32_t reversible;          /* EncodeParams.Lossless: 1 = 5/3, 0 = 9/7  encoder.go:31 */
    int32_t htj2k;               /* EncodeParams.HTJ2KMode: quant scale 1 instead of 64, no <<6  encoder.go:97,2312-2315,3294 */
    int32_t mct_mode;            /* j2k_mct_mode */
    double mct_matrix[J2K_MAX_COMPONENTS * J2K_MAX_COMPONENTS]; /* MCTMatrix, row-major C x C (custom modes) encoder.go:76 */
    int32_t mct_has_offsets;     /* MCTOffsets present and len == components  encoder.go:481 */
    int32_t mct_offsets[J2K_MAX_COMPONENTS];
    int32_t n_bindings;          /* MCTBindings, already in application order (encoder.go:532-537) */
    j2k_mct_binding bindings[J2K_MAX_BINDINGS];
    int32_t n_steps;             /* 9/7 only: 3*num_levels+1 runtime steps in QCD order, or 0 = plain rounding (encoder.go:2266-2273) */
    double steps[J2K_MAX_BANDS]; /* OpenJPEGRuntimeQuantizationSteps(...) values (quantization.go:140-154); <= 0 = plain rounding for that band (encoder.go:2320-2321) */
    int32_t fuse_t1_shift;       /* 1 = also apply the classic-EBCOT `<<6` of encodeCodeBlock (encoder.go:3294-3300); only meaningful when reversible && !htj2k */
    int32_t reserved[7];
} j2k_fwd_params;

/* Inverse (decode-side) parameters. */
typedef struct j2k_inv_params {
    /* SIZ geometry (jpeg2000/tile_assembler.go:33-56, jpeg2000/t2/tile_decoder.go:269-294) */
    int32_t xsiz, ysiz, xosiz, yosiz, xtsiz, ytsiz, xtosiz, ytosiz;
    int32_t components;          /* Csiz; XRsiz = YRsiz = 1 is the only supported sampling (tile_assembler.go:152-159) */
    int32_t bit_depth;           /* component 0 precision  decoder.go:143-144 */
    int32_t is_signed;
    int32_t num_levels;          /* COD NumberOfDecompositionLevels  t2/tile_decoder.go:352 */
    int32_t reversible;          /* COD Transformation: 1 = 5/3, 0 = 9/7  t2/tile_decoder.go:893-916 */
    int32_t htj2k;               /* COD code-block style bit 0x40  decoder.go:588; dequant by step instead of 0.5*step  t2/tile_decoder.go:974-977 */
    int32_t n_steps;             /* 9/7 only: number of decoded steps (<= 3L+1; bands beyond it are only converted, t2/tile_decoder.go:1022-1028); 0 = QCD style 0 (no dequantization, :909-911) */
    double steps[J2K_MAX_BANDS]; /* decodeQuantizationSteps(...) values WITHOUT the 0.5 factor (t2/tile_decoder.go:995-1046); <= 0 = band left unscaled (:971-973) */
    int32_t mct_mode;            /* NONE / RCT / ICT / CUSTOM_FLOAT / BINDINGS */
    double mct_matrix[J2K_MAX_COMPONENTS * J2K_MAX_COMPONENTS]; /* mctInverse, row-major C x C  decoder.go:696-711 */
    int32_t mct_has_offsets;
    int32_t mct_offsets[J2K_MAX_COMPONENTS];   /* added after the matrix  decoder.go:712-722 */
    int32_t n_bindings;
    j2k_mct_binding bindings[J2K_MAX_BINDINGS];
    int32_t fuse_t1_halve;       /* 1 = also apply the truncating `/2` of normalizeOpenJPEGReversibleT1Coefficients (t2/tile_decoder.go:989-993); only meaningful when reversible && !htj2k */
    int32_t reserved[7];
} j2k_inv_params;

/* CUDA-event timing of the most recent synchronous call on this context (milliseconds). */
typedef struct j2k_timing {
    float h2d_ms;
    float kernel_ms;
    float d2h_ms;
    float total_ms;
    int32_t kernel_launches;     /* kernels launched by that call */
    int32_t reserved;
} j2k_timing;

typedef struct j2k_ctx j2k_ctx;

/* ---------------------------------------------------------------- lifetime */

/* Create a context over `n_devices` CUDA devices (`devices == NULL` -> device 0..n-1,
 * n_devices == 0 -> the device named by env J2K_B200_DEVICE, default 0).
 * Replaces nothing in the reference (it has no device); called once from the
 * shim's init(), next to RegisterJPEG2000LosslessCodec (jpeg2000/lossless/codec.go:306-322). */
int j2k_init(j2k_ctx** ctx, const int* devices, int n_devices);
void j2k_shutdown(j2k_ctx* ctx);
/* Message of the most recent failing call on `ctx`, whichever thread made it (never NULL; the pointer stays valid on the
 * calling thread until its next j2k_last_error call).  ctx == NULL: the calling thread's last message (j2k_init and the
 * context-free helpers).  A goroutine may change OS threads between a failing cgo call and this one - the message is
 * kept per context for exactly that reason; wrap call + fetch in one Go function (integration/go/j2kb200.go). */
const char* j2k_last_error(j2k_ctx* ctx);
/* The same, copied into a caller buffer (at most cap - 1 bytes + NUL); returns the full message length. */
size_t j2k_last_error_copy(j2k_ctx* ctx, char* buf, size_t cap);
int j2k_abi_version(void);
int j2k_device_count(const j2k_ctx* ctx);
/* Failure detection (SURVEY 5 "a failed GPU is removed from the round-robin"; the reference has no counterpart, its errors
 * are plain Go `error`s, encoder.go:183-189).  When a CUDA call of a host-batch entry point fails and the device does not
 * answer a synchronise any more, its slot is marked failed: the blocking calls re-run that device's frame block on the
 * remaining devices and return J2K_OK if they succeed (the event is kept in j2k_last_error), asynchronous submissions
 * return the error, and every later call shards over the devices that are left.  1 = slot failed, 0 = in use, -1 = bad slot. */
int j2k_device_failed(const j2k_ctx* ctx, int slot);
/* CUDA devices visible to the process: `j2k_init(&ctx, NULL, j2k_visible_devices())` builds a context over all of them. */
int j2k_visible_devices(void);
/* Total number of CUDA kernels this context has launched (bench.py's gpu_launches). */
int64_t j2k_launch_count(const j2k_ctx* ctx);
int j2k_last_timing(j2k_ctx* ctx, j2k_timing* out);
/* Per-launch CUDA-event timing of plan runs (bench.py's roofline leg; off by default).  While
 * enabled every kernel of a run is bracketed by events on the launching stream.
 * j2k_get_profile synchronises and returns the launches of the most recent run: ms[i] is the
 * duration, levels[i] the DWT level the kernel computed (1 = finest) or 0 for a pointwise kernel. */
int j2k_set_profiling(j2k_ctx* ctx, int enabled);
int j2k_get_profile(j2k_ctx* ctx, float* ms, int32_t* levels, int max);

/* Library-owned pinned host memory (Go wraps it with unsafe.Slice). */
void* j2k_acquire_buffer(j2k_ctx* ctx, size_t nbytes);
void j2k_release_buffer(j2k_ctx* ctx, void* p);

/* ------------------------------------------------------------------- sizes */

/* expectedBytes of convertPixelData (jpeg2000/encoder.go:344). */
size_t j2k_fwd_pixel_bytes(const j2k_fwd_params* p);
/* width*height*components int32 values (all tiles, all components). */
size_t j2k_fwd_coeff_count(const j2k_fwd_params* p);
size_t j2k_inv_pixel_bytes(const j2k_inv_params* p);
size_t j2k_inv_coeff_count(const j2k_inv_params* p);
/* Tile grid of the encoder (jpeg2000/encoder.go:1966-1997) / decoder (tile_assembler.go:33-101):
 * writes x0,y0,x1,y1 (image-local, exclusive) of tile `idx`; returns the tile count. */
int j2k_fwd_tile_bounds(const j2k_fwd_params* p, int idx, int32_t bounds[4]);
int j2k_inv_tile_bounds(const j2k_inv_params* p, int idx, int32_t bounds[4]);

/* ----------------------------------------------------------------- forward */

/* Replaces, for one frame: Encoder.Encode's sample-domain head
 * (convertPixelData, applyDCLevelShift, MCT dispatch: jpeg2000/encoder.go:187-209,341-383,3698-3711)
 * plus, for every tile, Encoder.transformTile (jpeg2000/encoder.go:2213-2237) =
 * applyWaveletTransform / applyIrreversibleWaveletTransform / applyQuantizationBySubbandFloat
 * (:2187-2329).  `coeffs_out` receives what writeTile (:2099) hands to buildTilePacketEncoder. */
int j2k_forward(j2k_ctx* ctx, const j2k_fwd_params* p, const void* pixels, size_t nbytes,
                int32_t* coeffs_out, size_t ncoeffs);

/* Planar twin for Encoder.EncodeComponents([][]int32) (jpeg2000/encoder.go:221-273): skips convertPixelData. */
int j2k_forward_planar(j2k_ctx* ctx, const j2k_fwd_params* p, const int32_t* const* planes,
                       int32_t* coeffs_out, size_t ncoeffs);

/* The same with the planes in one buffer, component c at planes + c * plane_stride (samples): the cgo-callable form -
 * a Go [][]int32 holds Go pointers and cannot cross the boundary; the shim flattens it (integration/go/j2kb200.go). */
int j2k_forward_planar_flat(j2k_ctx* ctx, const j2k_fwd_params* p, const int32_t* planes, size_t plane_stride,
                            int32_t* coeffs_out, size_t ncoeffs);

/* Batched frames sharing one parameter set: the frame loops of the codec adapters
 * (jpeg2000/lossless/codec.go:246-261, jpeg2000/lossy/codec.go:149-176).  Frames are
 * contiguous: frame f at pixels + f*frame_stride_bytes, result at coeffs_out + f*coeff_count.
 * Frames are sharded over the context's devices in contiguous blocks; no collective. */
int j2k_forward_batch(j2k_ctx* ctx, const j2k_fwd_params* p, int nframes, const void* pixels,
                      size_t frame_stride_bytes, int32_t* coeffs_out);

/* Device-resident form: d_pixels / d_coeffs are device pointers on device index
 * `dev` of the context, work is enqueued on `cuda_stream` (a cudaStream_t, NULL = the
 * context's compute stream) and NOT synchronised. */
int j2k_forward_device(j2k_ctx* ctx, int dev, const j2k_fwd_params* p, int nframes,
                       const void* d_pixels, size_t frame_stride_bytes, int32_t* d_coeffs,
                       void* cuda_stream);

/* ----------------------------------------------------------------- inverse */

/* Replaces, for one frame: TileDecoder.applyIDWT for every tile-component
 * (jpeg2000/t2/tile_decoder.go:886-919 incl. applyDequantizationBySubbandFloat :925-987),
 * TileAssembler.AssembleTile (jpeg2000/tile_assembler.go:138-178),
 * Decoder.applyInverseTransforms + applyInverseDCLevelShift (jpeg2000/decoder.go:540-542,620-735,948-962)
 * and Decoder.GetPixelData (jpeg2000/decoder.go:777-944).
 * `planes_out` (optional, may be NULL) receives Decoder.GetImageData()
 * (jpeg2000/decoder.go:738-740): `components` planes of int32[H*W], unclamped. */
int j2k_inverse(j2k_ctx* ctx, const j2k_inv_params* p, const int32_t* coeffs_in, size_t ncoeffs,
                void* pixels_out, size_t nbytes, int32_t* planes_out);

int j2k_inverse_batch(j2k_ctx* ctx, const j2k_inv_params* p, int nframes, const int32_t* coeffs_in,
                      void* pixels_out, size_t frame_stride_bytes, int32_t* planes_out);

int j2k_inverse_device(j2k_ctx* ctx, int dev, const j2k_inv_params* p, int nframes,
                       const int32_t* d_coeffs, void* d_pixels, size_t frame_stride_bytes,
                       int32_t* d_planes, void* cuda_stream);

/* ------------------------------------------------------------ asynchronous */

/* Ticketed forms of the batch calls.  Buffers MUST come from j2k_acquire_buffer()
 * and stay untouched until j2k_wait() returns.  This is what lets the Go frame
 * loop run T1/T2 of frame i while the GPUs transform frames i+1... */
int64_t j2k_submit_forward(j2k_ctx* ctx, const j2k_fwd_params* p, int nframes, const void* pixels,
                           size_t frame_stride_bytes, int32_t* coeffs_out);
int64_t j2k_submit_inverse(j2k_ctx* ctx, const j2k_inv_params* p, int nframes, const int32_t* coeffs_in,
                           void* pixels_out, size_t frame_stride_bytes, int32_t* planes_out);
int j2k_wait(j2k_ctx* ctx, int64_t ticket);

/* ------------------------------------------- code-block interface */

/* One code-block of a tile-component plane, as Encoder.partitionIntoCodeBlocks
 * (jpeg2000/encoder.go:3215-3285) produces it from the sub-bands of
 * getSubbandsForResolution (:3059-3197), in the reference's order: resolution 0 (LL),
 * then per resolution HL, LH, HH; inside a sub-band row-major over (cby, cbx). */
typedef struct j2k_cblk {
    int32_t x0, y0;          /* codeBlockInfo.globalX0/globalY0: position in the tile-component plane */
    int32_t width, height;   /* actual size (clipped at the sub-band edge)                        */
    int32_t cbx, cby;        /* index inside the sub-band                                          */
    int32_t band, res;       /* 0 = LL, 1 = HL, 2 = LH, 3 = HH; resolution level                   */
    int64_t offset;          /* first sample of the block in the block-major plane (samples)       */
} j2k_cblk;

/* Block table of one tile-component plane of width x height samples; returns the number of
 * blocks (also with out == NULL).  The block-major plane has width*height samples, like the
 * Mallat plane it permutes.  Pure host arithmetic. */
int j2k_codeblock_layout(int width, int height, int num_levels, int cb_width, int cb_height,
                         j2k_cblk* out, int max_blocks);
/* Blocks per frame: all tiles (raster order), all components. */
size_t j2k_fwd_block_count(const j2k_fwd_params* p, int cb_width, int cb_height);
size_t j2k_inv_block_count(const j2k_inv_params* p, int cb_width, int cb_height);

/* j2k_forward_batch followed, on the device, by what buildTilePacketEncoder does before T1
 * (jpeg2000/encoder.go:2424-2431): sub-band extraction + partitionIntoCodeBlocks, the T1
 * scaling of encodeCodeBlock (:3294-3300: << 6 for classic lossless) and codeBlockNumBps
 * (:3349-3362 with calculateMaxBitplane :3643-3667).  `blocks_out`: per frame, per tile, per
 * component the block-major plane (each block contiguous, row-major); `numbps_out`: one
 * int32 per block in the same order (cblkNumbps).  ROI scaling stays in Go behind this call. */
int j2k_forward_blocks(j2k_ctx* ctx, const j2k_fwd_params* p, int cb_width, int cb_height, int nframes,
                       const void* pixels, size_t frame_stride_bytes, int32_t* blocks_out, int32_t* numbps_out);

/* Decode mirror: `blocks_in` holds the T1 output of every code-block (after the Go side's
 * inverse ROI scaling), block-major as above; the device performs TileDecoder.assembleSubbands
 * (jpeg2000/t2/tile_decoder.go:840-883) and then everything j2k_inverse_batch does.  The
 * classic 5/3 "/2" (tile_decoder.go:989-993) is applied when p->fuse_t1_halve is set. */
int j2k_inverse_blocks(j2k_ctx* ctx, const j2k_inv_params* p, int cb_width, int cb_height, int nframes,
                       const int32_t* blocks_in, void* pixels_out, size_t frame_stride_bytes, int32_t* planes_out);

/* The same with the decode-side MaxShift ROI of decodeCodeBlock (jpeg2000/t2/tile_decoder.go:726-730): the blocks arrive
 * as T1 leaves them and applyInverseMaxShift (:1113-1138) runs on the device while the blocks are scattered, before the
 * classic 5/3 "/2".  `roi_maxshift`: one shift per component (RGN marker, Srgn = 0), 0 = none; NULL = no ROI.  General
 * scaling (Srgn = 1, :735-741) depends on the ROI geometry per block and stays in Go. */
int j2k_inverse_blocks_roi(j2k_ctx* ctx, const j2k_inv_params* p, int cb_width, int cb_height, int nframes,
                           const int32_t* blocks_in, const int32_t* roi_maxshift, void* pixels_out, size_t frame_stride_bytes,
                           int32_t* planes_out);

/* The whole ROI tail of decodeCodeBlock (t2/tile_decoder.go:723-742) on the device, general scaling (RGN Srgn = 1) included:
 * after the MaxShift rule and the classic 5/3 "/2", a block that intersects the region has its samples - all of them, or
 * those its mask names - divided by 2^shift with Go's truncating division (applyInverseGeneralScaling /
 * applyInverseGeneralScalingMasked, :1082-1111).  The geometry stays in Go (ROIInfo.context / blockMask, :201-252); what
 * crosses the boundary is its result:
 *   block_scale_shift : nframes x j2k_inv_block_count() values, block order of j2k_codeblock_layout per tile-component:
 *                       the shift of a block with style == 1 && shiftVal > 0 && inside, else 0.  NULL = no general scaling.
 *   sample_mask       : optional, one byte per coefficient in block-major order (the layout of blocks_in): non-zero = the
 *                       sample lies in the region (blockMask).  NULL = whole blocks (the rectangle form).
 * Truncating divisions by powers of two commute, so the kernel applies the scaling while scattering, before the "/2". */
int j2k_inverse_blocks_roi_general(j2k_ctx* ctx, const j2k_inv_params* p, int cb_width, int cb_height, int nframes,
                                   const int32_t* blocks_in, const int32_t* roi_maxshift, const int32_t* block_scale_shift,
                                   const uint8_t* sample_mask, void* pixels_out, size_t frame_stride_bytes, int32_t* planes_out);

/* Device-resident halves of the two calls above (plane <-> block-major), for pipelines that
 * keep coefficients on the device; enqueued on `cuda_stream`, not synchronised.
 * The gather copies the coefficients AS THEY ARE (no shift) and counts cblkNumbps for them: with `p` exactly as it was
 * given to j2k_forward_device, numbps is right for 9/7 (6 fractional bits from the quantizer), HTJ2K (none) and 5/3 with
 * or without fuse_t1_shift (without it the plane holds plain integers and the caller applies the `<<6` of
 * encoder.go:3294-3300 to the block copies; the bit count of the integers is then the count T1 needs).
 * Planes that are not 16-byte aligned (a sliced device tensor) take scalar copies. */
int j2k_gather_blocks_device(j2k_ctx* ctx, int dev, const j2k_fwd_params* p, int cb_width, int cb_height, int nframes,
                             const int32_t* d_coeffs, int32_t* d_blocks, int32_t* d_numbps, void* cuda_stream);
int j2k_scatter_blocks_device(j2k_ctx* ctx, int dev, const j2k_inv_params* p, int cb_width, int cb_height, int nframes,
                              const int32_t* d_blocks, int32_t* d_coeffs, void* cuda_stream);
/* `roi_maxshift` is a HOST array (one shift per component, NULL = none), as in j2k_inverse_blocks_roi. */
int j2k_scatter_blocks_roi_device(j2k_ctx* ctx, int dev, const j2k_inv_params* p, int cb_width, int cb_height, int nframes,
                                  const int32_t* d_blocks, const int32_t* roi_maxshift, int32_t* d_coeffs, void* cuda_stream);
/* With general scaling: `d_block_scale_shift` (nframes x blocks, values 0..30) and `d_sample_mask` (optional, block-major bytes)
 * are DEVICE arrays, as j2k_inverse_blocks_roi_general describes them. */
int j2k_scatter_blocks_roi_general_device(j2k_ctx* ctx, int dev, const j2k_inv_params* p, int cb_width, int cb_height, int nframes,
                                          const int32_t* d_blocks, const int32_t* roi_maxshift, const int32_t* d_block_scale_shift,
                                          const uint8_t* d_sample_mask, int32_t* d_coeffs, void* cuda_stream);

/* ------------------------------- HTJ2K block decoding on the device (decode side) */

/* One HT code-block as T2 leaves it (cbInfo, jpeg2000/t2/tile_decoder.go:453-526): where its cleanup segment lies in the
 * caller's byte stream and the two facts HTDecoder.SetCodingContext receives (jpeg2000/htj2k/decoder.go:92-96, set at
 * t2/tile_decoder.go:601-609).  One record per (frame, block) in the order of the code-block interface: frame-major, then
 * tiles (raster), components, j2k_codeblock_layout order. */
typedef struct j2k_ht_cblk {
    uint64_t offset;        /* first byte of the segment in `bytes`                                                   */
    uint32_t length;        /* Lcup, bytes of the cleanup segment; 0 = block not included / no data: zero coefficients
                             * (HTDecoder.Decode, decoder.go:44-46; shouldDecode, t2/tile_decoder.go:672-689)          */
    uint8_t kmax;           /* bandNumbps of the block's sub-band (bandNumbpsFromQCD, t2/bitplane.go:22-61)           */
    uint8_t missing_msbs;   /* zero bit-planes from the packet header (htj2kMissingMSBs, t2/tile_decoder.go:691-699)  */
    uint16_t reserved;
} j2k_ht_cblk;

/* Per-block result codes in `status_out` (optional, one int32 per record).  A failing block decodes to zeros, which is what
 * TileDecoder.decodeCodeBlock substitutes when the block decoder returns an error (t2/tile_decoder.go:718-721); the call
 * itself still returns J2K_OK. */
#define J2K_HT_OK 0
#define J2K_HT_ERR_KMAX (-1)     /* "HTJ2K OpenJPH cleanup decoding requires band precision context" (decoder.go:48-50)        */
#define J2K_HT_ERR_SEGMENT (-2)  /* missing MSBs >= 30 or an invalid Scup locator (openjph_cleanup_decoder.go:122-127, decoder.go:63-65) */
#define J2K_HT_ERR_UQ (-3)       /* "U_q exceeds missing_msbs+2" (openjph_cleanup_decoder.go:292-294,333-335)                   */

/* HTDecoder.Decode (jpeg2000/htj2k/decoder.go:43-58 -> decodeOpenJPHCleanup, openjph_cleanup_decoder.go:115-161) for every
 * code-block of nframes frames: `blocks_out` receives the block-major planes j2k_inverse_blocks consumes.  HT code-blocks
 * hold at most 4096 samples (cb_width * cb_height <= 4096, ISO/IEC 15444-15); larger sizes return J2K_ERR_UNSUPPORTED.
 * Only the cleanup pass is decoded, as in the reference (SigProp / MagRef segments are ignored there too). */
int j2k_ht_decode_blocks(j2k_ctx* ctx, const j2k_inv_params* p, int cb_width, int cb_height, int nframes, const uint8_t* bytes,
                         size_t nbytes, const j2k_ht_cblk* cblks, int32_t* blocks_out, int32_t* status_out);

/* The whole decode tail on the device: HT block decoding written straight into the coefficient planes (assembleSubbands,
 * t2/tile_decoder.go:840-883, fused into the decoder's stores), then everything j2k_inverse_batch does.  Compressed bytes go
 * up, pixels come down.  p->htj2k should be set (HT blocks carry no T1 fixed point; no "/2", tile_decoder.go:732-734). */
int j2k_inverse_ht(j2k_ctx* ctx, const j2k_inv_params* p, int cb_width, int cb_height, int nframes, const uint8_t* bytes,
                   size_t nbytes, const j2k_ht_cblk* cblks, void* pixels_out, size_t frame_stride_bytes, int32_t* planes_out,
                   int32_t* status_out);

/* Ticketed form (j2k_wait); buffers from j2k_acquire_buffer, as for j2k_submit_inverse. */
int64_t j2k_submit_inverse_ht(j2k_ctx* ctx, const j2k_inv_params* p, int cb_width, int cb_height, int nframes, const uint8_t* bytes,
                              size_t nbytes, const j2k_ht_cblk* cblks, void* pixels_out, size_t frame_stride_bytes,
                              int32_t* planes_out, int32_t* status_out);

/* Device-resident form: `d_bytes`, `d_cblks` (nframes x blocks records), `d_out` and `d_status` (optional) are device
 * pointers; to_planes = 1 writes Mallat coefficient planes (input of j2k_inverse_device), 0 block-major planes.  Offsets and
 * lengths in device memory cannot be checked against the stream size: the caller guarantees offset + length <= nbytes. */
int j2k_ht_decode_device(j2k_ctx* ctx, int dev, const j2k_inv_params* p, int cb_width, int cb_height, int nframes,
                         const uint8_t* d_bytes, const j2k_ht_cblk* d_cblks, int32_t* d_out, int to_planes, int32_t* d_status,
                         void* cuda_stream);

/* The packed decode tables the device decoder indexes (which: 0 VLC initial row, 1 VLC other rows, 2 UVLC initial, 3 UVLC
 * other rows; vlc_tables.go:862-925, uvlc_tables.go:30-143); returns the entry count, copies them when out != NULL. */
int j2k_ht_table(int which, uint16_t* out);

/* ------------------------------- HTJ2K block encoding on the device (encode side) */

/* HTEncoder.Encode (jpeg2000/htj2k/encoder.go:54-68 -> encodeOpenJPHCleanup, openjph_cleanup_encoder.go:200-252) for every
 * code-block, behind j2k_forward_batch on the device: the coefficient planes never leave it; the blocks are read straight
 * from them (the sub-band extraction and partitionIntoCodeBlocks copies of buildTilePacketEncoder, encoder.go:2424-2431).
 * Byte-identical to the reference encoder (and to OpenJPH: htj2k/go_byte_parity_test.go).
 *   kmax      : components x (3 * num_levels + 1) band precisions, Encoder.bandNumbps per sub-band (index 0 = LL, then HL, LH,
 *               HH from the coarsest resolution); each 1..30 ("invalid HTJ2K Kmax", openjph_cleanup_encoder.go:201-203)
 *   bytes_out : the cleanup segments of all blocks, back to back (in block order when the context has one device; with several,
 *               sub-batch by sub-batch as the host collects them from the devices -- the records locate every segment);
 *               *nbytes_out receives their total size.
 *               When bytes_cap is too small the call fails with J2K_ERR_SIZE and *nbytes_out holds a size that was needed
 *               (j2k_ht_encode_bound gives a capacity that always suffices)
 *   cblks_out : one record per (frame, block) in the order of the code-block interface: offset / length of the block's
 *               segment (length 0: the block is empty, "return nil, nil" :218-220 -> not included in the packet), its Kmax and
 *               missing_msbs = Kmax - 1 (zeroBitPlanes of codeBlockPassLayout, encoder.go:3381-3388) -- what T2 needs.
 * p->htj2k should be set (no T1 fixed point, encoder.go:3293-3300).  The frames are sharded over the context's devices in
 * contiguous blocks like every host-batch call (no collective). */
int j2k_forward_ht(j2k_ctx* ctx, const j2k_fwd_params* p, int cb_width, int cb_height, int nframes, const void* pixels,
                   size_t frame_stride_bytes, const uint8_t* kmax, uint8_t* bytes_out, size_t bytes_cap, size_t* nbytes_out,
                   j2k_ht_cblk* cblks_out);
size_t j2k_ht_encode_bound(const j2k_fwd_params* p, int cb_width, int cb_height, int kmax_max, int nframes);

/* Device-resident form behind j2k_forward_device: `d_coeffs` (Mallat planes), `d_bytes` (capacity bytes_cap), `d_cblks`
 * (nframes x blocks records) and `d_offsets` (nframes x blocks + 1 values: the offset of every block; the last one is the
 * size of the stream, which is not written past bytes_cap) are device pointers; `kmax` is a host array as above. */
int j2k_ht_encode_device(j2k_ctx* ctx, int dev, const j2k_fwd_params* p, int cb_width, int cb_height, int nframes,
                         const int32_t* d_coeffs, const uint8_t* kmax, uint8_t* d_bytes, size_t bytes_cap, j2k_ht_cblk* d_cblks,
                         uint64_t* d_offsets, void* cuda_stream);

/* The encoder's VLC lookup (which: 0 initial quad row, 1 other rows; openjph_cleanup_encoder.go:440-470), 2048 entries. */
int j2k_ht_enc_table(int which, uint16_t* out);

/* -------------------------------------------- wavelet package API (in place) */

/* wavelet.ForwardMultilevelWithParity / InverseMultilevelWithParity
 * (jpeg2000/wavelet/dwt53.go:365-394,404-434): int32 plane, stride = width. */
int j2k_dwt53_forward(j2k_ctx* ctx, int32_t* data, int width, int height, int levels, int x0, int y0);
int j2k_dwt53_inverse(j2k_ctx* ctx, int32_t* data, int width, int height, int levels, int x0, int y0);
/* wavelet.ForwardMultilevel97Float32WithParity / InverseMultilevel97OpenJPEGWithParity
 * (jpeg2000/wavelet/dwt97.go:388-407,425-451): float32 plane, stride = width. */
int j2k_dwt97_forward(j2k_ctx* ctx, float* data, int width, int height, int levels, int x0, int y0);
int j2k_dwt97_inverse(j2k_ctx* ctx, float* data, int width, int height, int levels, int x0, int y0);
/* wavelet.ConvertFloat32ToInt32OpenJPEG (dwt97.go:473-503): round half to even. */
int j2k_convert_f32_to_i32(j2k_ctx* ctx, const float* in, int32_t* out, size_t n);
/* The float64 wrappers wavelet.ForwardMultilevel97WithParity / InverseMultilevel97WithParity
 * (dwt97.go:340-351,410-421): convert to float32, transform, convert back. */
int j2k_dwt97_forward_f64(j2k_ctx* ctx, double* data, int width, int height, int levels, int x0, int y0);
int j2k_dwt97_inverse_f64(j2k_ctx* ctx, double* data, int width, int height, int levels, int x0, int y0);
/* wavelet.ConvertFloat64ToInt32 (dwt97.go:515-526): truncate v +- 0.5 (half away from zero). */
int j2k_convert_f64_to_i32(j2k_ctx* ctx, const double* in, int32_t* out, size_t n);
/* wavelet.LLDimensionsWithParity (layout.go:11-33); LLDimensions is x0 = y0 = 0.  Host arithmetic. */
int j2k_ll_dimensions(int width, int height, int levels, int x0, int y0, int* ll_width, int* ll_height);

/* ----------------------------------------- colorspace package API (planar) */

/* colorspace.ApplyRCTToComponents / ApplyInverseRCTToComponents (colorspace/rct.go:26-49). */
int j2k_rct_forward(j2k_ctx* ctx, size_t n, const int32_t* r, const int32_t* g, const int32_t* b,
                    int32_t* y, int32_t* cb, int32_t* cr);
int j2k_rct_inverse(j2k_ctx* ctx, size_t n, const int32_t* y, const int32_t* cb, const int32_t* cr,
                    int32_t* r, int32_t* g, int32_t* b);
/* colorspace.ApplyICTToComponents / ApplyInverseICTToComponents (colorspace/ict.go:24-45): float64 + math.Round. */
int j2k_ict_forward(j2k_ctx* ctx, size_t n, const int32_t* r, const int32_t* g, const int32_t* b,
                    int32_t* y, int32_t* cb, int32_t* cr);
int j2k_ict_inverse(j2k_ctx* ctx, size_t n, const int32_t* y, const int32_t* cb, const int32_t* cr,
                    int32_t* r, int32_t* g, int32_t* b);
/* colorspace.ConvertRGBToYCbCr / ConvertYCbCrToRGB (colorspace/rgb.go:17-52): the same ICT on an interleaved
 * [R0,G0,B0,R1,...] image.  (ConvertComponentsRGBToYCbCr / ...YCbCrToRGB, rgb.go:100-123, are j2k_ict_forward / inverse.) */
int j2k_rgb_to_ycbcr(j2k_ctx* ctx, const int32_t* rgb, int width, int height, int32_t* y, int32_t* cb, int32_t* cr);
int j2k_ycbcr_to_rgb(j2k_ctx* ctx, const int32_t* y, const int32_t* cb, const int32_t* cr, int width, int height, int32_t* rgb);
/* colorspace.InterleaveComponents / DeinterleaveComponents (colorspace/rgb.go:54-98): planes [C][n] <-> [n][C]. */
int j2k_interleave_components(j2k_ctx* ctx, const int32_t* const* components, int n_components, size_t n_pixels, int32_t* out);
int j2k_deinterleave_components(j2k_ctx* ctx, const int32_t* data, size_t n_pixels, int n_components, int32_t* const* components_out);

/* -------------------------------------------------- quantization.go API */

/* QuantizeCoefficients / DequantizeCoefficients (jpeg2000/quantization.go:310-340):
 * RoundToEven(float64(c)/step), RoundToEven(float64(c)*step); step <= 0 copies. */
int j2k_quantize_coefficients(j2k_ctx* ctx, const int32_t* in, int32_t* out, size_t n, double step);
int j2k_dequantize_coefficients(j2k_ctx* ctx, const int32_t* in, int32_t* out, size_t n, double step);

/* Scalar step-table metadata (no device work).  These stay in Go in a real
 * integration; they are exported so that non-Go hosts
 * (bench.py, the Python mirror) can build the same tables.
 *   j2k_quant_openjpeg_params  = CalculateOpenJPEGQuantizationParams (quantization.go:212-236)
 *   j2k_quant_quality_params   = CalculateQuantizationParams         (quantization.go:180-208)
 *   j2k_quant_runtime_steps    = OpenJPEGRuntimeQuantizationSteps    (quantization.go:140-154)
 *   j2k_quant_decode_steps     = TileDecoder.decodeQuantizationSteps, style 2 (t2/tile_decoder.go:1018-1043)
 * `encoded`/`steps` hold 3*num_levels+1 entries.  Return the entry count or a negative status. */
int j2k_quant_openjpeg_params(int num_levels, int bit_depth, uint16_t* encoded, double* step_sizes);
int j2k_quant_quality_params(int quality, int num_levels, int bit_depth, uint16_t* encoded, double* step_sizes);
int j2k_quant_runtime_steps(const uint16_t* encoded, int n, int num_levels, int bit_depth, double* steps);
int j2k_quant_decode_steps(const uint16_t* encoded, int n, int num_levels, int bit_depth, int reversible,
                           double* steps);

#ifdef __cplusplus
}
#endif
#endif /* J2K_B200_H */
