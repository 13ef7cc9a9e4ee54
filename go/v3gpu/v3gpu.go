// Package v3gpu is the cgo binding of libgochem_b200.so (include/gochem_b200.h): device-resident batches of v3 matrices
// and the fused superposition / RMSD / per-atom-deviation passes over them.
//
// It follows the only FFI precedent of goChem, traj/xtc/xtc.go:30-38, 95-121, 173-183: a cgo preamble, an opaque C
// handle kept in a Go struct, int status returns (0 = ok), explicit Close plus a finalizer (traj/dcd/dcd.go:266-268).
//
// Build:  CGO_CFLAGS="-I<repo>/include" CGO_LDFLAGS="-L<repo>/gochem_b200/lib -lgochem_b200" go build -tags b200 ./...
//
// Handles are not goroutine-safe (one goroutine per handle, like the trajectory readers); distinct handles may be used
// from different goroutines at the same time -- every setting and scratch buffer lives in the handle.
package v3gpu

/*
#cgo LDFLAGS: -lgochem_b200
#include <stdlib.h>
#include <string.h>
#include "gochem_b200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"

	v3 "github.com/rmera/gochem/v3"
)

// Host layouts accepted by the staging calls (GCB_F64_AOS ...).
const (
	F64AoS = int(C.GCB_F64_AOS) // v3.Matrix.RawSlice, v3/gonum.go:93-95
	F32AoS = int(C.GCB_F32_AOS) // libxdrfile output, traj/xtc/libgoxtc.c:87-104
	F32SoA = int(C.GCB_F32_SOA) // DCD planes, traj/dcd/dcd.go:380-413
	I32AoS = int(C.GCB_I32_AOS) // STF fixed point, traj/stf/stf.go:324-345
)

// Error carries the library's status code (GCB_ERR_*) and message.
type Error struct {
	Code int
	Msg  string
}

func (e *Error) Error() string { return fmt.Sprintf("gochem_b200 (%d): %s", e.Code, e.Msg) }

// call runs one C entry point and, on failure, reads the message the library keeps per OS thread.  The goroutine is
// pinned to its thread in between: cgo calls may otherwise hop threads and read another call's message.
func call(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	rc := f()
	if rc == 0 {
		return nil
	}
	buf := make([]C.char, 2048)
	C.gcb_last_error(&buf[0], C.size_t(len(buf)))
	return &Error{Code: int(rc), Msg: C.GoString(&buf[0])}
}

// cints converts a Go index list ([]int, what chem.Super and align.RMSDTraj take) to the C ABI's 32-bit ints.
// A nil or empty list gives a nil pointer ("all atoms", handy.go:178-180).
func cints(idx []int) ([]C.int, *C.int) {
	if len(idx) == 0 {
		return nil, nil
	}
	out := make([]C.int, len(idx))
	for i, v := range idx {
		out[i] = C.int(v)
	}
	return out, &out[0]
}

func cbool(b bool) C.int {
	if b {
		return 1
	}
	return 0
}

// DeviceCount returns the number of CUDA devices the library sees (0 without a usable driver).
func DeviceCount() int {
	var n C.int
	if C.gcb_device_count(&n) != 0 {
		return 0
	}
	return int(n)
}

// ---------------------------------------------------------------------------------------------------------------------
// Pinned: C-owned page-locked staging memory (Go memory may not be retained across an asynchronous copy).
// ---------------------------------------------------------------------------------------------------------------------

type Pinned struct {
	p unsafe.Pointer
	n int
}

func NewPinned(bytes int) (*Pinned, error) {
	b := &Pinned{n: bytes}
	if err := call(func() C.int { return C.gcb_pinned_alloc(C.size_t(bytes), &b.p) }); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(b, func(b *Pinned) { b.Free() })
	return b, nil
}

func (b *Pinned) Free() {
	if b.p != nil {
		C.gcb_pinned_free(b.p)
		b.p = nil
	}
}

// Bytes exposes the buffer as a slice a reader can io.ReadFull into (raw DCD records, traj/dcd/dcd.go:349-434).
func (b *Pinned) Bytes() []byte { return unsafe.Slice((*byte)(b.p), b.n) }

// Float32s / Float64s view the same memory as coordinates.
func (b *Pinned) Float32s() []float32 { return unsafe.Slice((*float32)(b.p), b.n/4) }
func (b *Pinned) Float64s() []float64 { return unsafe.Slice((*float64)(b.p), b.n/8) }

// ---------------------------------------------------------------------------------------------------------------------
// Traj: the device-resident batch of v3 matrices (cap frames of natoms x 3, float64 SoA planes in HBM).
// ---------------------------------------------------------------------------------------------------------------------

type Traj struct {
	h      *C.gcb_traj
	natoms int
	arena  [2]*Pinned // staging for Upload, allocated on first use
}

func NewTraj(dev, natoms int, capFrames int64) (*Traj, error) {
	t := &Traj{natoms: natoms}
	if err := call(func() C.int { return C.gcb_traj_create(C.int(dev), C.int(natoms), C.long(capFrames), &t.h) }); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(t, func(t *Traj) { t.Close() })
	return t, nil
}

// Close frees the device memory; explicit like xtc.XTCObj.Close (traj/xtc/xtc.go:173-183).
func (t *Traj) Close() {
	if t.h != nil {
		C.gcb_traj_destroy(t.h)
		t.h = nil
	}
	for i := range t.arena {
		if t.arena[i] != nil {
			t.arena[i].Free()
			t.arena[i] = nil
		}
	}
}

func (t *Traj) NAtoms() int      { return t.natoms }
func (t *Traj) NFrames() int64   { return int64(C.gcb_traj_nframes(t.h)) }
func (t *Traj) Capacity() int64  { return int64(C.gcb_traj_capacity(t.h)) }
func (t *Traj) StageBytes() int  { return int(C.gcb_traj_stage_bytes(t.h)) }
func (t *Traj) SetNFrames(n int64) error {
	return call(func() C.int { return C.gcb_traj_set_nframes(t.h, C.long(n)) })
}

// Upload copies frames (each a *v3.Matrix, row-major N x 3) into slots [first, first+len(frames)).  The rows of the
// matrices are packed back to back into a pinned arena -- one copy() per frame, no cgo call -- and every full arena goes
// to the device with ONE gcb_traj_upload_async; the two arenas alternate, so packing overlaps the previous DMA.
func (t *Traj) Upload(first int64, frames []*v3.Matrix) error {
	per := 3 * t.natoms // float64 per frame
	room := t.StageBytes() / (8 * per)
	if room < 1 {
		return &Error{Code: int(C.GCB_ERR_ARG), Msg: "a frame does not fit the staging slot"}
	}
	for i := range t.arena {
		if t.arena[i] == nil {
			a, err := NewPinned(room * per * 8)
			if err != nil {
				return err
			}
			t.arena[i] = a
		}
	}
	slot := 0
	for done := 0; done < len(frames); slot ^= 1 {
		n := len(frames) - done
		if n > room {
			n = room
		}
		if err := t.Wait(slot); err != nil { // the DMA out of this arena (two batches ago) has finished
			return err
		}
		dst := t.arena[slot].Float64s()
		for k := 0; k < n; k++ {
			f := frames[done+k]
			if f.NVecs() != t.natoms {
				return &Error{Code: int(C.GCB_ERR_SHAPE), Msg: fmt.Sprintf("frame %d has %d atoms, the trajectory %d", done+k, f.NVecs(), t.natoms)}
			}
			copy(dst[k*per:(k+1)*per], f.RawSlice())
		}
		if err := t.UploadPinned(first+int64(done), t.arena[slot], int64(n), F64AoS, 1.0, slot); err != nil {
			return err
		}
		done += n
	}
	if err := t.Wait(0); err != nil {
		return err
	}
	return t.Wait(1)
}

// UploadPinned queues nframes frames that a reader decoded straight into a pinned buffer (layout F32SoA for DCD planes,
// F32AoS with scale 10 for libxdrfile's nanometres, I32AoS with scale 100 for STF integers) on copy stream slot (0 or 1).
// The buffer may be refilled after Wait(slot).
func (t *Traj) UploadPinned(first int64, buf *Pinned, nframes int64, layout int, scale float64, slot int) error {
	return call(func() C.int {
		return C.gcb_traj_upload_async(t.h, C.long(first), buf.p, C.long(nframes), C.int(layout), C.double(scale), C.int(slot))
	})
}

// UploadRecords queues nframes raw DCD frames -- Fortran record markers and unit-cell block included -- as they were
// read from the file: frame f is the recordBytes bytes at buf + f*recordBytes with the float32 planes at offX/offY/offZ.
func (t *Traj) UploadRecords(first int64, buf *Pinned, nframes, recordBytes, offX, offY, offZ int64, scale float64, slot int) error {
	return call(func() C.int {
		return C.gcb_traj_upload_records_async(t.h, C.long(first), buf.p, C.long(nframes), C.long(recordBytes), C.long(offX),
			C.long(offY), C.long(offZ), C.double(scale), C.int(slot))
	})
}

func (t *Traj) Wait(slot int) error {
	return call(func() C.int { return C.gcb_traj_wait(t.h, C.int(slot)) })
}

func (t *Traj) Sync() error { return call(func() C.int { return C.gcb_sync(t.h) }) }

// Download copies frames [first, first+len(out)) back into the given matrices (what chem.Super leaves in test).
func (t *Traj) Download(first int64, out []*v3.Matrix) error {
	per := 3 * t.natoms
	buf := make([]float64, per*len(out))
	if len(out) == 0 {
		return nil
	}
	err := call(func() C.int { return C.gcb_traj_download(t.h, C.long(first), C.long(len(out)), (*C.double)(&buf[0])) })
	if err != nil {
		return err
	}
	for k, m := range out {
		copy(m.RawSlice(), buf[k*per:(k+1)*per])
	}
	return nil
}

// Fit holds what chem.RotatorTranslatorToSuper returns per frame (geometric.go:206-207) plus chem.RMSD after the fit.
type Fit struct {
	RMSD   []float64 // F
	Rot    []float64 // 9F, row-major, right-multiplying row vectors
	Trans1 []float64 // 3F, minus the centroid of the test subset
	Trans2 []float64 // 3F
}

// SuperRMSD = chem.Super(frame, ref, idxTest, idxRef) + chem.RMSD for frames [first, first+n): handy.go:175-214,
// geometric.go:127-208, 232-288.  writeBack replaces every frame by its superposed coordinates (Super's in-place semantics).
func (t *Traj) SuperRMSD(first, n int64, ref *v3.Matrix, idxTest, idxRef []int, writeBack bool) (*Fit, error) {
	keepT, it := cints(idxTest)
	keepR, ir := cints(idxRef)
	if it != nil && ir == nil {
		ir = it
	}
	f := &Fit{RMSD: make([]float64, n), Rot: make([]float64, 9*n), Trans1: make([]float64, 3*n), Trans2: make([]float64, 3*n)}
	if n == 0 {
		return f, nil
	}
	raw := ref.RawSlice()
	err := call(func() C.int {
		return C.gcb_super_rmsd(t.h, C.long(first), C.long(n), (*C.double)(&raw[0]), C.int(ref.NVecs()), it, ir, C.int(len(idxTest)),
			cbool(writeBack), (*C.double)(&f.RMSD[0]), (*C.double)(&f.Rot[0]), (*C.double)(&f.Trans1[0]), (*C.double)(&f.Trans2[0]))
	})
	runtime.KeepAlive(keepT)
	runtime.KeepAlive(keepR)
	if err != nil {
		return nil, err
	}
	return f, nil
}

// RMSD = chem.RMSD / chem.MemRMSD without superposition for frames [first, first+n) (geometric.go:232-288).
func (t *Traj) RMSD(first, n int64, ref *v3.Matrix, idxTest, idxRef []int) ([]float64, error) {
	keepT, it := cints(idxTest)
	keepR, ir := cints(idxRef)
	if it != nil && ir == nil {
		ir = it
	}
	out := make([]float64, n)
	if n == 0 {
		return out, nil
	}
	raw := ref.RawSlice()
	err := call(func() C.int {
		return C.gcb_rmsd(t.h, C.long(first), C.long(n), (*C.double)(&raw[0]), C.int(ref.NVecs()), it, ir, C.int(len(idxTest)), (*C.double)(&out[0]))
	})
	runtime.KeepAlive(keepT)
	runtime.KeepAlive(keepR)
	return out, err
}

// PerAtomDist = chem.MemPerAtomRMSD for one resident frame (geometric.go:294-321): out[k] = |frame[idxTest[k]] - ref[idxRef[k]]|.
func (t *Traj) PerAtomDist(frame int64, ref *v3.Matrix, idxTest, idxRef []int) ([]float64, error) {
	keepT, it := cints(idxTest)
	keepR, ir := cints(idxRef)
	if it != nil && ir == nil {
		ir = it
	}
	n := len(idxTest)
	if n == 0 {
		n = t.natoms
	}
	out := make([]float64, n)
	raw := ref.RawSlice()
	err := call(func() C.int {
		return C.gcb_peratom_dist(t.h, C.long(frame), (*C.double)(&raw[0]), C.int(ref.NVecs()), it, ir, C.int(len(idxTest)), (*C.double)(&out[0]))
	})
	runtime.KeepAlive(keepT)
	runtime.KeepAlive(keepR)
	return out, err
}

// PerAtomDevSum = the body of align.RMSDTraj's loop over frames [first, first+n) (align/lovo.go:265-281, 328-340): Super
// on idx, then the distance of EVERY atom to ref, summed over the frames (not divided).  comm != nil combines the ranks' sums.
func (t *Traj) PerAtomDevSum(first, n int64, ref *v3.Matrix, idx []int, writeBack bool, comm *Comm) ([]float64, error) {
	keep, ip := cints(idx)
	sum := make([]float64, t.natoms)
	raw := ref.RawSlice()
	err := call(func() C.int {
		return C.gcb_peratom_dev_sum(t.h, C.long(first), C.long(n), (*C.double)(&raw[0]), ip, C.int(len(idx)), cbool(writeBack), comm.handle(),
			(*C.double)(&sum[0]))
	})
	runtime.KeepAlive(keep)
	return sum, err
}

// PerAtomDevSumAtoms returns the same sums for the atoms of want only (sum[k] belongs to want[k]): what align.LOVO keeps of
// RMSDTraj's result after rmsdFilter (align/lovo.go:232, 519-528).  Frames are left untouched.
func (t *Traj) PerAtomDevSumAtoms(first, n int64, ref *v3.Matrix, idx, want []int, comm *Comm) ([]float64, error) {
	keepI, ip := cints(idx)
	keepW, wp := cints(want)
	sum := make([]float64, len(want))
	if len(want) == 0 {
		return sum, nil
	}
	raw := ref.RawSlice()
	err := call(func() C.int {
		return C.gcb_peratom_dev_sum_atoms(t.h, C.long(first), C.long(n), (*C.double)(&raw[0]), ip, C.int(len(idx)), wp, C.int(len(want)),
			comm.handle(), (*C.double)(&sum[0]))
	})
	runtime.KeepAlive(keepI)
	runtime.KeepAlive(keepW)
	return sum, err
}

// LOVOResult is what the library's own LOVO iteration leaves: LOVOReturn.N and .Iterations, indexesold (its first N entries
// are LOVOReturn.Natoms) and the candidates with their mean deviations in the order of the last sort.
type LOVOResult struct {
	N, Iterations, Disagreement int
	IndexesOld, IDs             []int
	Vals                        []float64
}

// LOVOResident runs align.LOVO's fixed-point iteration (align/lovo.go:212-252) over the resident frames with the bookkeeping of
// lovo.go:229-252 on the device (gcb_lovo_resident): 64 bytes come back per iteration instead of the per-atom means.  taken is
// false when the library leaves the case to the caller (candidates not ascending, more than 16384 of them, small frames ...):
// run the loop with PerAtomDevSum then -- the answers are the same bit for bit.
func (t *Traj) LOVOResident(first, n int64, ref *v3.Matrix, cand []int, totalFrames int, lessThanRMSD float64, minimumN, maxIter int,
	comm *Comm) (res *LOVOResult, taken bool, err error) {
	if len(cand) == 0 {
		return nil, false, nil
	}
	keep, cp := cints(cand)
	old, ids := make([]C.int, len(cand)), make([]C.int, len(cand))
	vals := make([]float64, len(cand))
	var nOut, iters, dis C.int
	raw := ref.RawSlice()
	var rc C.int
	err = call(func() C.int {
		rc = C.gcb_lovo_resident(t.h, C.long(first), C.long(n), (*C.double)(&raw[0]), cp, C.int(len(cand)), C.double(totalFrames),
			C.double(lessThanRMSD), C.int(minimumN), C.int(maxIter), comm.handle(), &nOut, &iters, &dis, &old[0], &ids[0],
			(*C.double)(&vals[0]))
		if rc == 1 { // not taken: not an error
			return 0
		}
		return rc
	})
	runtime.KeepAlive(keep)
	if err != nil || rc == 1 {
		return nil, false, err
	}
	res = &LOVOResult{N: int(nOut), Iterations: int(iters), Disagreement: int(dis), IndexesOld: make([]int, len(cand)),
		IDs: make([]int, len(cand)), Vals: vals}
	for i := range old {
		res.IndexesOld[i], res.IDs[i] = int(old[i]), int(ids[i])
	}
	return res, true, nil
}

// AllPairsRMSD fills rows [row0, row0+nrows) of the nframes x nframes matrix D[i][j] = chem.RMSD(frame_i, frame_j)
// (fit = false; superpose first) or chem.RMSD(chem.Super(frame_i, frame_j), frame_j) (fit = true), one fp64 tensor-core GEMM.
func (t *Traj) AllPairsRMSD(idx []int, row0, nrows int64, fit bool) ([]float64, error) {
	keep, ip := cints(idx)
	n := t.NFrames()
	out := make([]float64, nrows*n)
	err := call(func() C.int {
		if fit {
			return C.gcb_allpairs_rmsd_fit(t.h, 0, C.long(n), ip, C.int(len(idx)), C.long(row0), C.long(nrows), (*C.double)(&out[0]), nil, nil)
		}
		return C.gcb_allpairs_rmsd(t.h, 0, C.long(n), ip, C.int(len(idx)), C.long(row0), C.long(nrows), (*C.double)(&out[0]), nil, nil)
	})
	runtime.KeepAlive(keep)
	return out, err
}

// Moments = chem.CenterOfMass (geometric.go:609-631) and chem.MomentTensor (geometric.go:705-733) of every frame.
func (t *Traj) Moments(first, n int64, mass []float64, idx []int) (center, moment []float64, err error) {
	keep, ip := cints(idx)
	center, moment = make([]float64, 3*n), make([]float64, 9*n)
	var mp *C.double
	if len(mass) > 0 {
		mp = (*C.double)(&mass[0])
	}
	err = call(func() C.int {
		return C.gcb_moments(t.h, C.long(first), C.long(n), mp, ip, C.int(len(idx)), (*C.double)(&center[0]), (*C.double)(&moment[0]))
	})
	runtime.KeepAlive(keep)
	return center, moment, err
}

// Settings of this handle (kernel choice and shape); the defaults suit every shape measured.
func (t *Traj) SetKernelVariant(v int) error {
	return call(func() C.int { return C.gcb_traj_set_kernel_variant(t.h, C.int(v)) })
}
func (t *Traj) SetExplicitRMSD(on bool) error {
	return call(func() C.int { return C.gcb_traj_set_rmsd_mode(t.h, cbool(on)) })
}

// ---------------------------------------------------------------------------------------------------------------------
// Comm: one process per GPU, NCCL over NVLink.  Rank 0 creates the id, the launcher broadcasts its 128 bytes.
// ---------------------------------------------------------------------------------------------------------------------

type Comm struct{ h *C.gcb_comm }

func (c *Comm) handle() *C.gcb_comm {
	if c == nil {
		return nil
	}
	return c.h
}

func UniqueID() ([128]byte, error) {
	var id [128]byte
	err := call(func() C.int { return C.gcb_comm_unique_id((*C.char)(unsafe.Pointer(&id[0]))) })
	return id, err
}

func NewComm(dev, nranks, rank int, id [128]byte) (*Comm, error) {
	c := &Comm{}
	err := call(func() C.int {
		return C.gcb_comm_init(C.int(dev), C.int(nranks), C.int(rank), (*C.char)(unsafe.Pointer(&id[0])), &c.h)
	})
	if err != nil {
		return nil, err
	}
	return c, nil
}

func (c *Comm) Close() {
	if c != nil && c.h != nil {
		C.gcb_comm_destroy(c.h)
		c.h = nil
	}
}

// AllReduceSum adds vec over the ranks in rank order (bit-identical on every rank).
func (c *Comm) AllReduceSum(vec []float64) error {
	if len(vec) == 0 {
		return nil
	}
	return call(func() C.int { return C.gcb_allreduce_sum_f64(c.h, (*C.double)(&vec[0]), C.long(len(vec))) })
}
