// Package gpu binds libj2kb200.so (include/j2k_b200.h) into go-dicom-codec.
//
// Drop this file at jpeg2000/gpu/j2kb200.go of github.com/cocosip/go-dicom-codec and apply the three call-site changes of
// INTEGRATION.md section 2.  It was written against the C header; Go is not installed in the image this repository is
// built in, so it has not been compiled there.  Every C entry point it calls is exercised by tests/ through the Python
// mirror (go-dicom-codec_b200/j2kb200) and by the plain C caller tests/c/abi_smoke.c.
package gpu

/*
#cgo CFLAGS:  -I${SRCDIR}/../../third_party/j2k_b200/include
#cgo LDFLAGS: -L${SRCDIR}/../../third_party/j2k_b200/lib -lj2kb200 -lcudart
#include <stdlib.h>
#include "j2k_b200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sync"
	"unsafe"
)

var (
	once sync.Once
	ctx  *C.j2k_ctx
	ierr error
)

// Init creates the process-wide context over EVERY visible device (frames and tiles of a batch are sharded over the
// GPUs in contiguous blocks; there is no collective).  devices == nil -> devices 0 .. j2k_visible_devices()-1.  Safe to
// call from any goroutine, any number of times.
func Init() error {
	once.Do(func() {
		runtime.LockOSThread() // j2k_init's message (no context yet) is per thread: fetch it on the thread that failed
		defer runtime.UnlockOSThread()
		n := C.j2k_visible_devices()
		if n <= 0 {
			ierr = fmt.Errorf("j2k_b200: no CUDA device visible; the GPU path has no CPU fallback")
			return
		}
		if rc := C.j2k_init(&ctx, nil, n); rc != 0 {
			ierr = fmt.Errorf("j2k_b200: init failed (%d): %s", int(rc), C.GoString(C.j2k_last_error(nil)))
		}
	})
	return ierr
}

// InitDevices is Init over an explicit device list (one context per process; the first call wins).
func InitDevices(devices []int) error {
	once.Do(func() {
		runtime.LockOSThread()
		defer runtime.UnlockOSThread()
		ids := make([]C.int, len(devices))
		for i, d := range devices {
			ids[i] = C.int(d)
		}
		var p *C.int
		if len(ids) > 0 {
			p = &ids[0]
		}
		if rc := C.j2k_init(&ctx, p, C.int(len(ids))); rc != 0 {
			ierr = fmt.Errorf("j2k_b200: init failed (%d): %s", int(rc), C.GoString(C.j2k_last_error(nil)))
		}
	})
	return ierr
}

// DeviceCount reports how many GPUs the context shards batches over.
func DeviceCount() int { return int(C.j2k_device_count(ctx)) }

// DeviceFailed reports whether the GPU in `slot` was lost (a CUDA call failed and the device no longer answers) and has been
// removed from the round-robin: blocking calls re-run its frame block on the remaining GPUs, later calls skip it.
func DeviceFailed(slot int) bool { return C.j2k_device_failed(ctx, C.int(slot)) == 1 }

// Shutdown releases the context (streams, device scratch, pinned staging).
func Shutdown() {
	if ctx != nil {
		C.j2k_shutdown(ctx)
		ctx = nil
	}
}

// lastErr turns a negative status into an error.  The library keeps the text of a failure per CONTEXT (not per OS
// thread), so it is still there when the goroutine has been moved to another thread between the failing cgo call and
// this one; j2k_last_error_copy writes into a Go-owned buffer, so no C pointer outlives the call.
func lastErr(rc C.int) error {
	var buf [512]C.char
	C.j2k_last_error_copy(ctx, &buf[0], C.size_t(len(buf)))
	return fmt.Errorf("j2k_b200 error %d: %s", int(rc), C.GoString(&buf[0]))
}

// MCT modes (j2k_mct_mode).
const (
	MCTNone        = int(C.J2K_MCT_NONE)
	MCTRCT         = int(C.J2K_MCT_RCT)
	MCTICT         = int(C.J2K_MCT_ICT)
	MCTCustomInt   = int(C.J2K_MCT_CUSTOM_INT)
	MCTCustomQ13   = int(C.J2K_MCT_CUSTOM_Q13)
	MCTCustomFloat = int(C.J2K_MCT_CUSTOM_FLOAT)
	MCTBindings    = int(C.J2K_MCT_BINDINGS)
)

// Binding mirrors jpeg2000.MCTBindingParams (encoder.go:111-121) / the decoder's mctBinding (decoder.go:630-694).
type Binding struct {
	ComponentIDs []int
	ElementType  int       // encode: 0 = integer matrix, else Q13; decode: 0 = integer, else float64
	Matrix       []float64 // row-major n x n, nil = identity (encode) / skipped (decode)
	Offsets      []int32
}

func (b *Binding) c() C.j2k_mct_binding {
	var c C.j2k_mct_binding
	c.n_components = C.int32_t(len(b.ComponentIDs))
	for i, id := range b.ComponentIDs {
		c.component_ids[i] = C.int32_t(id)
	}
	c.element_type = C.int32_t(b.ElementType)
	if b.Matrix != nil {
		c.has_matrix = 1
		for i, v := range b.Matrix {
			c.matrix[i] = C.double(v)
		}
	}
	if b.Offsets != nil {
		c.has_offsets = 1
		for i, v := range b.Offsets {
			c.offsets[i] = C.int32_t(v)
		}
	}
	return c
}

// FwdParams mirrors the fields of jpeg2000.EncodeParams the sample path reads (encoder.go:17-98).
type FwdParams struct {
	Width, Height, Components, BitDepth int
	IsSigned                            bool
	TileWidth, TileHeight, NumLevels    int
	Lossless, HTJ2K                     bool
	MCTMode                             int
	MCTMatrix                           []float64 // custom modes: row-major C x C (EncodeParams.MCTMatrix)
	MCTOffsets                          []int32
	Bindings                            []Binding
	Steps                               []float64 // OpenJPEGRuntimeQuantizationSteps(...), nil when Lossless
	FuseT1Shift                         bool      // classic EBCOT lossless without ROI only
}

func (p *FwdParams) c() C.j2k_fwd_params {
	var c C.j2k_fwd_params
	c.width, c.height, c.components = C.int32_t(p.Width), C.int32_t(p.Height), C.int32_t(p.Components)
	c.bit_depth = C.int32_t(p.BitDepth)
	if p.IsSigned {
		c.is_signed = 1
	}
	c.tile_width, c.tile_height, c.num_levels = C.int32_t(p.TileWidth), C.int32_t(p.TileHeight), C.int32_t(p.NumLevels)
	if p.Lossless {
		c.reversible = 1
	}
	if p.HTJ2K {
		c.htj2k = 1
	}
	c.mct_mode = C.int32_t(p.MCTMode)
	for i, v := range p.MCTMatrix {
		c.mct_matrix[i] = C.double(v)
	}
	if len(p.MCTOffsets) == p.Components && p.Components > 0 {
		c.mct_has_offsets = 1
		for i, v := range p.MCTOffsets {
			c.mct_offsets[i] = C.int32_t(v)
		}
	}
	c.n_bindings = C.int32_t(len(p.Bindings))
	for i := range p.Bindings {
		c.bindings[i] = p.Bindings[i].c()
	}
	c.n_steps = C.int32_t(len(p.Steps))
	for i, s := range p.Steps {
		c.steps[i] = C.double(s)
	}
	if p.FuseT1Shift {
		c.fuse_t1_shift = 1
	}
	return c
}

// InvParams carries what TileDecoder / Decoder read from SIZ, COD and QCD (t2/tile_decoder.go:269-294,886-987,
// decoder.go:143-144,588,620-735).
type InvParams struct {
	Xsiz, Ysiz, XOsiz, YOsiz, XTsiz, YTsiz, XTOsiz, YTOsiz int
	Components, BitDepth                                  int
	IsSigned                                              bool
	NumLevels                                             int
	Reversible, HTJ2K                                     bool      // COD transformation == 1; code-block style bit 0x40
	Steps                                                 []float64 // raw decodeQuantizationSteps(...) values, WITHOUT the 0.5 factor (the library applies 0.5*step for classic, step for HTJ2K)
	MCTMode                                               int
	MCTMatrix                                             []float64
	MCTOffsets                                            []int32
	Bindings                                              []Binding
	FuseT1Halve                                           bool
}

func (p *InvParams) c() C.j2k_inv_params {
	var c C.j2k_inv_params
	c.xsiz, c.ysiz, c.xosiz, c.yosiz = C.int32_t(p.Xsiz), C.int32_t(p.Ysiz), C.int32_t(p.XOsiz), C.int32_t(p.YOsiz)
	c.xtsiz, c.ytsiz, c.xtosiz, c.ytosiz = C.int32_t(p.XTsiz), C.int32_t(p.YTsiz), C.int32_t(p.XTOsiz), C.int32_t(p.YTOsiz)
	c.components, c.bit_depth, c.num_levels = C.int32_t(p.Components), C.int32_t(p.BitDepth), C.int32_t(p.NumLevels)
	if p.IsSigned {
		c.is_signed = 1
	}
	if p.Reversible {
		c.reversible = 1
	}
	if p.HTJ2K {
		c.htj2k = 1
	}
	c.n_steps = C.int32_t(len(p.Steps))
	for i, s := range p.Steps {
		c.steps[i] = C.double(s)
	}
	c.mct_mode = C.int32_t(p.MCTMode)
	for i, v := range p.MCTMatrix {
		c.mct_matrix[i] = C.double(v)
	}
	if len(p.MCTOffsets) == p.Components && p.Components > 0 {
		c.mct_has_offsets = 1
		for i, v := range p.MCTOffsets {
			c.mct_offsets[i] = C.int32_t(v)
		}
	}
	c.n_bindings = C.int32_t(len(p.Bindings))
	for i := range p.Bindings {
		c.bindings[i] = p.Bindings[i].c()
	}
	if p.FuseT1Halve {
		c.fuse_t1_halve = 1
	}
	return c
}

// CoeffCount / PixelBytes size the caller's buffers.
func (p *FwdParams) CoeffCount() int { cp := p.c(); return int(C.j2k_fwd_coeff_count(&cp)) }
func (p *FwdParams) PixelBytes() int { cp := p.c(); return int(C.j2k_fwd_pixel_bytes(&cp)) }
func (p *InvParams) CoeffCount() int { cp := p.c(); return int(C.j2k_inv_coeff_count(&cp)) }
func (p *InvParams) PixelBytes() int { cp := p.c(); return int(C.j2k_inv_pixel_bytes(&cp)) }

// Forward: one frame, interleaved pixel bytes in, all tiles' coefficient planes out (tile-major, component-major,
// row-major, stride = tile width) - what transformTile returns for every tile.  A Go slice is pageable memory: the C
// side moves it through its pinned staging ring (host threads fill 4 MB chunks while the copy engines drain them) and
// has copied everything out of / into the slices when it returns: no Go pointer is retained (cgo rule).
func Forward(p *FwdParams, pixels []byte, coeffs []int32) error {
	cp := p.c()
	rc := C.j2k_forward(ctx, &cp, unsafe.Pointer(&pixels[0]), C.size_t(len(pixels)),
		(*C.int32_t)(unsafe.Pointer(&coeffs[0])), C.size_t(len(coeffs)))
	if rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// ForwardPlanar is the binding of Encoder.EncodeComponents([][]int32) (encoder.go:221-273): component planes in, no byte
// conversion.  A [][]int32 holds Go pointers to Go memory and must not be passed to C, so the planes are flattened into
// one slice (component c at c*Width*Height) - when the caller's planes already are consecutive windows of one backing
// array (as convertPixelData's successor can allocate them) the copy is skipped.
func ForwardPlanar(p *FwdParams, planes [][]int32, coeffs []int32) error {
	hw := p.Width * p.Height
	if len(planes) != p.Components {
		return fmt.Errorf("expected %d components, got %d", p.Components, len(planes)) // encoder.go:229-231
	}
	for i, pl := range planes {
		if len(pl) != hw {
			return fmt.Errorf("component %d: expected %d pixels, got %d", i, hw, len(pl)) // encoder.go:234-238
		}
	}
	flat := planes[0][:hw:hw]
	contiguous := true
	for c := 1; c < len(planes) && contiguous; c++ {
		contiguous = uintptr(unsafe.Pointer(&planes[c][0])) == uintptr(unsafe.Pointer(&planes[0][0]))+uintptr(4*c*hw)
	}
	if contiguous && len(planes) > 1 {
		flat = unsafe.Slice(&planes[0][0], hw*len(planes))
	} else if len(planes) > 1 {
		flat = make([]int32, hw*len(planes))
		for c, pl := range planes {
			copy(flat[c*hw:], pl)
		}
	}
	cp := p.c()
	rc := C.j2k_forward_planar_flat(ctx, &cp, (*C.int32_t)(unsafe.Pointer(&flat[0])), C.size_t(hw),
		(*C.int32_t)(unsafe.Pointer(&coeffs[0])), C.size_t(len(coeffs)))
	if rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// ForwardBatch: nframes contiguous frames sharing one parameter set (the adapters' frame loops).
func ForwardBatch(p *FwdParams, nframes int, pixels []byte, frameStride int, coeffs []int32) error {
	cp := p.c()
	rc := C.j2k_forward_batch(ctx, &cp, C.int(nframes), unsafe.Pointer(&pixels[0]), C.size_t(frameStride),
		(*C.int32_t)(unsafe.Pointer(&coeffs[0])))
	if rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// Inverse: every tile-component's coefficient plane in (after assembleSubbands), packed pixel bytes out; planes, when
// non-nil, receives Decoder.GetImageData() (Components planes of Width*Height int32).
func Inverse(p *InvParams, coeffs []int32, pixels []byte, planes []int32) error {
	cp := p.c()
	var pl *C.int32_t
	if planes != nil {
		pl = (*C.int32_t)(unsafe.Pointer(&planes[0]))
	}
	rc := C.j2k_inverse(ctx, &cp, (*C.int32_t)(unsafe.Pointer(&coeffs[0])), C.size_t(len(coeffs)),
		unsafe.Pointer(&pixels[0]), C.size_t(len(pixels)), pl)
	if rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// InverseBatch: nframes frames, coefficient planes and pixel frames contiguous.
func InverseBatch(p *InvParams, nframes int, coeffs []int32, pixels []byte, frameStride int) error {
	cp := p.c()
	rc := C.j2k_inverse_batch(ctx, &cp, C.int(nframes), (*C.int32_t)(unsafe.Pointer(&coeffs[0])),
		unsafe.Pointer(&pixels[0]), C.size_t(frameStride), nil)
	if rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// ---- asynchronous: pinned buffers owned by the library, tickets (INTEGRATION.md 2.3)

// PinnedBytes / PinnedInt32 wrap a page-locked buffer from j2k_acquire_buffer; Release gives it back.
func PinnedBytes(n int) ([]byte, error) {
	p := C.j2k_acquire_buffer(ctx, C.size_t(n))
	if p == nil {
		return nil, lastErr(C.int(C.J2K_ERR_NOMEM))
	}
	return unsafe.Slice((*byte)(p), n), nil
}

func PinnedInt32(n int) ([]int32, error) {
	p := C.j2k_acquire_buffer(ctx, C.size_t(n)*4)
	if p == nil {
		return nil, lastErr(C.int(C.J2K_ERR_NOMEM))
	}
	return unsafe.Slice((*int32)(p), n), nil
}

func ReleaseBytes(b []byte)   { C.j2k_release_buffer(ctx, unsafe.Pointer(&b[0])) }
func ReleaseInt32(b []int32)  { C.j2k_release_buffer(ctx, unsafe.Pointer(&b[0])) }

// Ticket identifies a submitted batch; it completes when that batch's last device-to-host copy has landed.
type Ticket int64

// SubmitForward enqueues nframes frames and returns at once.  pixels and coeffs MUST be pinned buffers from this
// package and stay untouched until Wait returns.
func SubmitForward(p *FwdParams, nframes int, pixels []byte, frameStride int, coeffs []int32) (Ticket, error) {
	cp := p.c()
	t := C.j2k_submit_forward(ctx, &cp, C.int(nframes), unsafe.Pointer(&pixels[0]), C.size_t(frameStride),
		(*C.int32_t)(unsafe.Pointer(&coeffs[0])))
	if t < 0 {
		return 0, lastErr(C.int(t))
	}
	return Ticket(t), nil
}

func SubmitInverse(p *InvParams, nframes int, coeffs []int32, pixels []byte, frameStride int) (Ticket, error) {
	cp := p.c()
	t := C.j2k_submit_inverse(ctx, &cp, C.int(nframes), (*C.int32_t)(unsafe.Pointer(&coeffs[0])),
		unsafe.Pointer(&pixels[0]), C.size_t(frameStride), nil)
	if t < 0 {
		return 0, lastErr(C.int(t))
	}
	return Ticket(t), nil
}

func Wait(t Ticket) error {
	if rc := C.j2k_wait(ctx, C.int64_t(t)); rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// ---- code-block interface (INTEGRATION.md 2.4)

// CodeBlock mirrors codeBlockInfo's geometry (encoder.go:3199-3212) plus the block's offset in the block-major plane.
type CodeBlock struct {
	X0, Y0, Width, Height, CBX, CBY, Band, Res int
	Offset                                     int64
}

// CodeBlockLayout returns the blocks of one tile-component plane in the order buildTilePacketEncoder walks them.
func CodeBlockLayout(width, height, numLevels, cbWidth, cbHeight int) ([]CodeBlock, error) {
	n := C.j2k_codeblock_layout(C.int(width), C.int(height), C.int(numLevels), C.int(cbWidth), C.int(cbHeight), nil, 0)
	if n < 0 {
		return nil, lastErr(n)
	}
	if n == 0 {
		return nil, nil
	}
	raw := make([]C.j2k_cblk, int(n))
	C.j2k_codeblock_layout(C.int(width), C.int(height), C.int(numLevels), C.int(cbWidth), C.int(cbHeight), &raw[0], n)
	out := make([]CodeBlock, int(n))
	for i, b := range raw {
		out[i] = CodeBlock{int(b.x0), int(b.y0), int(b.width), int(b.height), int(b.cbx), int(b.cby), int(b.band), int(b.res), int64(b.offset)}
	}
	return out, nil
}

// ForwardBlocks: like ForwardBatch, but the coefficients arrive partitioned into code-blocks (each contiguous, the T1
// shift applied) together with cblkNumbps per block.
func ForwardBlocks(p *FwdParams, cbWidth, cbHeight, nframes int, pixels []byte, frameStride int, blocks []int32, numbps []int32) error {
	cp := p.c()
	rc := C.j2k_forward_blocks(ctx, &cp, C.int(cbWidth), C.int(cbHeight), C.int(nframes), unsafe.Pointer(&pixels[0]),
		C.size_t(frameStride), (*C.int32_t)(unsafe.Pointer(&blocks[0])), (*C.int32_t)(unsafe.Pointer(&numbps[0])))
	if rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// InverseBlocksROI: InverseBlocks for codestreams with an RGN marker of style 0 (MaxShift): blocks holds the T1 output
// untouched and roiMaxShift[c] the shift of component c (0 = none); applyInverseMaxShift (t2/tile_decoder.go:1113-1138)
// runs on the device while the blocks are scattered.  General scaling (Srgn = 1) stays in decodeCodeBlock.
func InverseBlocksROI(p *InvParams, cbWidth, cbHeight, nframes int, blocks []int32, roiMaxShift []int32, pixels []byte, frameStride int) error {
	cp := p.c()
	var roi *C.int32_t
	if len(roiMaxShift) > 0 {
		roi = (*C.int32_t)(unsafe.Pointer(&roiMaxShift[0]))
	}
	rc := C.j2k_inverse_blocks_roi(ctx, &cp, C.int(cbWidth), C.int(cbHeight), C.int(nframes), (*C.int32_t)(unsafe.Pointer(&blocks[0])),
		roi, unsafe.Pointer(&pixels[0]), C.size_t(frameStride), nil)
	if rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// InverseBlocks: the decode mirror; blocks holds every code-block's T1 output at CodeBlock.Offset.
func InverseBlocks(p *InvParams, cbWidth, cbHeight, nframes int, blocks []int32, pixels []byte, frameStride int) error {
	cp := p.c()
	rc := C.j2k_inverse_blocks(ctx, &cp, C.int(cbWidth), C.int(cbHeight), C.int(nframes), (*C.int32_t)(unsafe.Pointer(&blocks[0])),
		unsafe.Pointer(&pixels[0]), C.size_t(frameStride), nil)
	if rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// HTBlock mirrors j2k_ht_cblk: one HT code-block as T2 leaves it (cbInfo, t2/tile_decoder.go:453-526) plus the coding
// context HTDecoder.SetCodingContext receives (htj2k/decoder.go:92-96; set at t2/tile_decoder.go:601-609).
type HTBlock struct {
	Offset      uint64 // first byte of the cleanup segment in the stream handed to InverseHT
	Length      uint32 // Lcup; 0 = block not included (shouldDecode, t2/tile_decoder.go:672-689): zero coefficients
	Kmax        uint8  // bandNumbpsFromQCD (t2/bitplane.go:22-61)
	MissingMSBs uint8  // htj2kMissingMSBs (t2/tile_decoder.go:691-699)
	_           uint16
}

// InverseHT: the whole HTJ2K decode tail on the device.  stream holds the cleanup segments of every code-block of nframes
// frames (T2 already parsed the packets), blocks one record per (frame, block) in CodeBlockLayout order; the device runs
// HTDecoder.Decode (htj2k/decoder.go:43-58) for every block, assembleSubbands (t2/tile_decoder.go:840-883) and everything
// InverseBatch does.  status (optional, one entry per record) receives 0 or the block decoder's error code; a failing
// block reads as zeros, as decodeCodeBlock substitutes (t2/tile_decoder.go:718-721).
func InverseHT(p *InvParams, cbWidth, cbHeight, nframes int, stream []byte, blocks []HTBlock, pixels []byte, frameStride int, status []int32) error {
	cp := p.c()
	var st *C.int32_t
	if len(status) > 0 {
		st = (*C.int32_t)(unsafe.Pointer(&status[0]))
	}
	var sp *C.uint8_t
	if len(stream) > 0 {
		sp = (*C.uint8_t)(unsafe.Pointer(&stream[0]))
	}
	rc := C.j2k_inverse_ht(ctx, &cp, C.int(cbWidth), C.int(cbHeight), C.int(nframes), sp, C.size_t(len(stream)),
		(*C.j2k_ht_cblk)(unsafe.Pointer(&blocks[0])), unsafe.Pointer(&pixels[0]), C.size_t(frameStride), nil, st)
	if rc != 0 {
		return lastErr(rc)
	}
	return nil
}

// ForwardHT: the whole HTJ2K encode front on the device.  pixels holds nframes frames; kmax the band precisions
// (Encoder.bandNumbps, encoder.go:1687-1694: components x (3*levels+1), index 0 = LL, then HL, LH, HH from the coarsest
// resolution).  The device runs everything ForwardBatch does and then HTEncoder.Encode (htj2k/encoder.go:54-68) for every
// code-block, reading the blocks straight from the coefficient planes.  stream receives the cleanup segments back to back
// (cap(stream) must hold them: C.j2k_ht_encode_bound gives a size that always does); blocks one record per (frame, block) in
// CodeBlockLayout order: Length == 0 is an empty block (HTEncoder returns nil: not included in the packet), MissingMSBs is
// the zeroBitPlanes codeBlockPassLayout would compute (encoder.go:3381-3388).  Returns the number of stream bytes.
func ForwardHT(p *FwdParams, cbWidth, cbHeight, nframes int, pixels []byte, frameStride int, kmax []uint8, stream []byte, blocks []HTBlock) (int, error) {
	cp := p.c()
	var n C.size_t
	rc := C.j2k_forward_ht(ctx, &cp, C.int(cbWidth), C.int(cbHeight), C.int(nframes), unsafe.Pointer(&pixels[0]), C.size_t(frameStride),
		(*C.uint8_t)(unsafe.Pointer(&kmax[0])), (*C.uint8_t)(unsafe.Pointer(&stream[0])), C.size_t(cap(stream)), &n,
		(*C.j2k_ht_cblk)(unsafe.Pointer(&blocks[0])))
	if rc != 0 {
		return int(n), lastErr(rc)
	}
	return int(n), nil
}
