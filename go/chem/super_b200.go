//go:build b200

// Drop-in replacements (package chem, build tag b200) for the superposition / RMSD entry points of goChem, bound to
// libgochem_b200.so through package v3gpu.  Signatures are the reference's, byte for byte:
//
//	Super                     handy.go:175
//	RotatorTranslatorToSuper  geometric.go:127
//	RMSD                      geometric.go:232
//	MemRMSD                   geometric.go:258
//	MemPerAtomRMSD            geometric.go:294
//
// go/apply.sh moves the CPU bodies of those five functions into *_cpu.go files tagged !b200, so `go build` without the
// tag is the unmodified reference and `go build -tags b200` is this file.  There is no CPU fallback under the tag: without
// a CUDA device every call returns the library's error.
//
// A single structure is a batch of one on the device (latency-bound, ~30 us per call); trajectories should go through
// align.RMSDTraj / align.LOVO (go/align/lovo_b200.go) or v3gpu.Traj directly, which keep the frames resident.
package chem

import (
	"fmt"
	"sync"

	v3 "github.com/rmera/gochem/v3"
	"github.com/rmera/gochem/v3gpu"
)

// one-frame device handles, pooled per atom count: each call takes its own, so the functions stay reentrant on distinct
// matrices, as align.RMSDTraj needs (align/lovo.go:290-295, 329)
var (
	oneFrameMu    sync.Mutex
	oneFramePools = map[int]*sync.Pool{}
)

// B200Device is the CUDA device the single-structure calls run on.
var B200Device = 0

func oneFrame(natoms int) (*v3gpu.Traj, func(), error) {
	oneFrameMu.Lock()
	p, ok := oneFramePools[natoms]
	if !ok {
		p = &sync.Pool{}
		oneFramePools[natoms] = p
	}
	oneFrameMu.Unlock()
	if h, ok := p.Get().(*v3gpu.Traj); ok && h != nil {
		return h, func() { p.Put(h) }, nil
	}
	h, err := v3gpu.NewTraj(B200Device, natoms, 1)
	if err != nil {
		return nil, nil, err
	}
	return h, func() { p.Put(h) }, nil
}

func identity(n int) []int {
	out := make([]int, n)
	for i := range out {
		out[i] = i
	}
	return out
}

// fitLists is the index dispatch shared by Super, MemRMSD and MemPerAtomRMSD (handy.go:178-201, geometric.go:261-283,
// 295-312): which rows of test pair with which rows of templa.  nil, nil means "all atoms of both".
// who prefixes the error text ("chem.Super", "chem.memRMSD", "chem.memMSD").
func fitLists(test, templa *v3.Matrix, indexes [][]int, who, illFormed string) (it, ir []int, err error) {
	switch {
	case len(indexes) == 0 || indexes[0] == nil || len(indexes[0]) == 0:
		if test.NVecs() != templa.NVecs() {
			return nil, nil, fmt.Errorf("%s", illFormed)
		}
		return nil, nil, nil
	case len(indexes) == 1:
		n := len(indexes[0])
		if test.NVecs() > n {
			if templa.NVecs() != n {
				return nil, nil, fmt.Errorf("%s", illFormed)
			}
			return indexes[0], identity(n), nil
		} else if templa.NVecs() > n {
			// the reference leaves ctest nil here and dereferences it (handy.go:186-188): a panic there, an error here
			return nil, nil, fmt.Errorf("%s: a single index list that applies to templa leaves test unset (nil dereference in the reference)", who)
		}
		return nil, nil, fmt.Errorf("%s: Indexes don't match molecules", who)
	default:
		if len(indexes[0]) != len(indexes[1]) {
			return nil, nil, fmt.Errorf("%s", illFormed)
		}
		return indexes[0], indexes[1], nil
	}
}

// Super determines the best rotation and translations to superimpose test on templa, considering the atoms in indexes
// (same dispatch as the reference), and transforms ALL of test in place: test = (test + trans1) R + trans2.
func Super(test, templa *v3.Matrix, indexes ...[]int) (*v3.Matrix, error) {
	it, ir, err := fitLists(test, templa, indexes, "chem.Super", "chem.Super: Ill formed coordinates for Superposition")
	if err != nil {
		return nil, err
	}
	h, done, err := oneFrame(test.NVecs())
	if err != nil {
		return nil, errDecorate(err, "Super")
	}
	defer done()
	if err = h.Upload(0, []*v3.Matrix{test}); err != nil {
		return nil, errDecorate(err, "Super")
	}
	if _, err = h.SuperRMSD(0, 1, templa, it, ir, true); err != nil {
		return nil, errDecorate(errDecorate(err, "RotatorTranslatorToSuper"), "Super")
	}
	if err = h.Download(0, []*v3.Matrix{test}); err != nil {
		return nil, errDecorate(err, "Super")
	}
	return test, nil
}

// RotatorTranslatorToSuper returns the transformed copy of test, the rotation (3x3, right-multiplying row vectors), minus
// the centroid of test (1x3) and the final translation (1x3), as geometric.go:127-208 does.
func RotatorTranslatorToSuper(test, templa *v3.Matrix) (*v3.Matrix, *v3.Matrix, *v3.Matrix, *v3.Matrix, error) {
	tmr, tmc := templa.Dims()
	tsr, tsc := test.Dims()
	if tmr != tsr || tmc != 3 || tsc != 3 {
		return nil, nil, nil, nil, CError{"goChem: Ill-formed matrices", []string{"RotatorTranslatorToSuper"}}
	}
	h, done, err := oneFrame(tsr)
	if err != nil {
		return nil, nil, nil, nil, errDecorate(err, "RotatorTranslatorToSuper")
	}
	defer done()
	if err = h.Upload(0, []*v3.Matrix{test}); err != nil {
		return nil, nil, nil, nil, errDecorate(err, "RotatorTranslatorToSuper")
	}
	fit, err := h.SuperRMSD(0, 1, templa, nil, nil, true)
	if err != nil {
		return nil, nil, nil, nil, errDecorate(err, "RotatorTranslatorToSuper")
	}
	transformed := v3.Zeros(tsr)
	if err = h.Download(0, []*v3.Matrix{transformed}); err != nil {
		return nil, nil, nil, nil, errDecorate(err, "RotatorTranslatorToSuper")
	}
	rotation, err := v3.NewMatrix(fit.Rot)
	if err != nil {
		return nil, nil, nil, nil, errDecorate(err, "RotatorTranslatorToSuper")
	}
	trans1, _ := v3.NewMatrix(fit.Trans1)
	trans2, _ := v3.NewMatrix(fit.Trans2)
	return transformed, rotation, trans1, trans2, nil
}

// RMSD calculates the RMSD between test and templa over the atoms in indexes, without superposition.
func RMSD(test, templa *v3.Matrix, indexes ...[]int) (float64, error) {
	var L int
	if len(indexes) == 0 || indexes[0] == nil || len(indexes[0]) == 0 {
		L = test.NVecs()
	} else {
		L = len(indexes[0])
	}
	return MemRMSD(test, templa, v3.Zeros(L), indexes...)
}

// MemRMSD is RMSD with the caller's scratch matrix; tmp is only checked for its size here (the difference is formed on the device).
func MemRMSD(test, templa, tmp *v3.Matrix, indexes ...[]int) (float64, error) {
	const ill = "chem.memRMSD: Ill formed matrices for memRMSD calculation"
	it, ir, err := fitLists(test, templa, indexes, "chem.memRMSD", ill)
	if err != nil {
		return -1, err
	}
	n := len(it)
	if n == 0 {
		n = test.NVecs()
	}
	if tmp.NVecs() != n {
		return -1, fmt.Errorf(ill)
	}
	h, done, err := oneFrame(test.NVecs())
	if err != nil {
		return -1, errDecorate(err, "MemRMSD")
	}
	defer done()
	if err = h.Upload(0, []*v3.Matrix{test}); err != nil {
		return -1, errDecorate(err, "MemRMSD")
	}
	out, err := h.RMSD(0, 1, templa, it, ir)
	if err != nil {
		return -1, errDecorate(err, "MemRMSD")
	}
	return out[0], nil
}

// MemPerAtomRMSD returns the distance of every considered atom of test to its partner in templa (not squared, not a mean:
// geometric.go:314-319).  ctest, ctempla and tmp are the reference's scratch matrices; only tmp's size is checked.
func MemPerAtomRMSD(test, templa, ctest, ctempla, tmp *v3.Matrix, indexes ...[]int) ([]float64, error) {
	it, ir, err := fitLists(test, templa, indexes, "chem.memMSD", "chem.memMSD: Ill formed matrices for memMSD calculation")
	if err != nil {
		return nil, err
	}
	n := len(it)
	if n == 0 {
		n = test.NVecs()
	}
	if tmp.NVecs() != n {
		return nil, fmt.Errorf("chem.memMSD: Ill formed matrices for memMSD calculation: cest: %d  ctempla: %d tmp: %d", n, n, tmp.NVecs())
	}
	h, done, err := oneFrame(test.NVecs())
	if err != nil {
		return nil, errDecorate(err, "MemPerAtomRMSD")
	}
	defer done()
	if err = h.Upload(0, []*v3.Matrix{test}); err != nil {
		return nil, errDecorate(err, "MemPerAtomRMSD")
	}
	return h.PerAtomDist(0, templa, it, ir)
}
