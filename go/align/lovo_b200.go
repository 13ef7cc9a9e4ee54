//go:build b200

// Drop-in replacements (package align, build tag b200) for the trajectory loop of goChem's LOVO alignment, bound to
// libgochem_b200.so through package v3gpu.  Signatures are the reference's, byte for byte:
//
//	RMSDTraj  align/lovo.go:286
//	LOVO      align/lovo.go:198
//
// go/apply.sh moves the CPU bodies of LOVO, concproc, RMSDTraj and the rmsdandcoords type into lovo_cpu.go (tag !b200);
// everything else of lovo.go -- Options, LOVOReturn, opentraj, mobileLimit, the rMSD sorter, sameElementsInt, res2atoms ... --
// is shared and untouched, so the index sets come out of the reference's own bookkeeping.
//
// What changes: frames are not handed to o.cpus goroutines one chunk at a time (lovo.go:315-340) but read into a
// device-resident trajectory (float64 SoA planes in HBM) and every RMSDTraj pass -- superposition of each frame on the index
// subset plus the distance of every atom to the reference -- is one fused kernel launch per window of frames.  LOVO reads
// the file ONCE and keeps the frames on the device across its iterations (the reference re-opens and re-decodes the file in
// every iteration, lovo.go:222).  The chunk semantics are kept exactly: frames are consumed in chunks of o.cpus, Begin and
// Skip discard chunks with the reference's arithmetic (including the chunk that follows a discarded one, lovo.go:316-322),
// a chunk the reader cannot fill is dropped, and the means are divided by chunksread*chunksize (lovo.go:348-351).
package align

import (
	"fmt"
	"log"

	chem "github.com/rmera/gochem"
	"github.com/rmera/gochem/traj/dcd"
	v3 "github.com/rmera/gochem/v3"
	"github.com/rmera/gochem/v3gpu"
)

// B200Device is the CUDA device the alignment runs on; B200WindowBytes bounds one device window of frames
// (trajectories larger than the device memory are processed window by window, the per-window sums added on the host).
var (
	B200Device      = 0
	B200WindowBytes = int64(8) << 30
)

// the reference's loop has no bound (lovo.go:221); the library's wants one
const b200MaxIterations = 1 << 20

// deviceFrames is the selected part of a trajectory, resident on the device as one or more windows.
type deviceFrames struct {
	windows []*v3gpu.Traj
	counts  []int64
	natoms  int
	total   int // chunksread * chunksize
}

func (d *deviceFrames) Close() {
	for _, w := range d.windows {
		w.Close()
	}
	d.windows = nil
}

// loadSelected reads traj exactly as RMSDTraj's loop does (lovo.go:306-326) and uploads every selected chunk.
func loadSelected(natoms int, traj chem.ConcTraj, o *Options) (*deviceFrames, error) {
	chunksize := o.cpus
	begin := o.begin / chunksize
	skip := o.skip / chunksize
	chunk := make([]*v3.Matrix, chunksize)
	for i := range chunk {
		chunk[i] = v3.Zeros(natoms)
	}
	perWindow := B200WindowBytes / int64(24*natoms)
	perWindow -= perWindow % int64(chunksize)
	if perWindow < int64(chunksize) {
		perWindow = int64(chunksize)
	}
	d := &deviceFrames{natoms: natoms}
	var cur *v3gpu.Traj
	var filled int64
	var err error
	chunksread := 0
	for read := 0; ; read++ {
		if read < begin || read%(skip+1) != 0 {
			discard := make([]*v3.Matrix, chunksize) // all nil: read and throw away (interfaces.go:51-64)
			if _, err = traj.NextConc(discard); err != nil {
				break
			}
		}
		var chans []chan *v3.Matrix
		if chans, err = traj.NextConc(chunk); err != nil {
			break
		}
		for _, c := range chans { // the readers fill chunk[i] and signal through the channel
			<-c
		}
		chunksread++
		if cur == nil || filled+int64(chunksize) > perWindow {
			if cur != nil {
				d.counts = append(d.counts, filled)
			}
			if cur, err = v3gpu.NewTraj(B200Device, natoms, perWindow); err != nil {
				d.Close()
				return nil, err
			}
			d.windows = append(d.windows, cur)
			filled = 0
		}
		if err = cur.Upload(filled, chunk); err != nil {
			d.Close()
			return nil, err
		}
		filled += int64(chunksize)
	}
	if cur != nil {
		d.counts = append(d.counts, filled)
	}
	if err == nil {
		log.Println("Finished reading trajectory without a EOF signal. This shouldn't happen.")
	} else if _, ok := err.(chem.LastFrameError); !ok {
		d.Close()
		return nil, fmt.Errorf("Error while reading trajectory, read %d sets of frames. Error: %s", chunksread, err.Error())
	}
	d.total = chunksread * chunksize
	return d, nil
}

// meanDeviations is one RMSDTraj pass over resident frames: per-atom distance to ref after the fit on indexes, averaged.
func (d *deviceFrames) meanDeviations(ref *v3.Matrix, indexes []int, writeBack bool) ([]float64, error) {
	RMSD := make([]float64, d.natoms)
	for w, t := range d.windows {
		part, err := t.PerAtomDevSum(0, d.counts[w], ref, indexes, writeBack, nil)
		if err != nil {
			return nil, err
		}
		for i, v := range part {
			RMSD[i] += v
		}
	}
	totalframes := float64(d.total)
	for i, v := range RMSD {
		RMSD[i] = v / totalframes
	}
	return RMSD, nil
}

// writeAligned writes the (already superposed) resident frames as a DCD file, like the reference's wtraj.WNext calls
// (lovo.go:296-305, 337-339).  A file that cannot be opened is reported and skipped, as there.
func (d *deviceFrames) writeAligned(name string) {
	wtraj, err := dcd.NewWriter(name, d.natoms)
	if err != nil {
		log.Printf("Couldn't open %s for writing. Will not write aligned trajectory", name)
		return
	}
	defer wtraj.Close()
	buf := []*v3.Matrix{v3.Zeros(d.natoms)}
	for w, t := range d.windows {
		for f := int64(0); f < d.counts[w]; f++ {
			if t.Download(f, buf) != nil {
				return
			}
			wtraj.WNext(buf[0])
		}
	}
}

// RMSDTraj returns the RMSD for all atoms in a structure, averaged over the trajectory traj.
// Frames are read in sets of cpus at the time: The read starts at the set begin/cpus and we skip a set every skip/cpu
// (the numbers are, of course, rounded). The RMSDs and the total number of frames read are returned, together with an error or nil.
func RMSDTraj(mol chem.Atomer, ref *v3.Matrix, traj chem.ConcTraj, indexes []int, o *Options) ([]float64, int, error) {
	d, err := loadSelected(mol.Len(), traj, o)
	if err != nil {
		return nil, -1, err
	}
	defer d.Close()
	RMSD, err := d.meanDeviations(ref, indexes, o.writeTraj != "")
	if err != nil {
		return nil, -1, err
	}
	if o.writeTraj != "" {
		d.writeAligned(o.writeTraj)
	}
	return RMSD, d.total, nil
}

// LOVO returns slices with the indexes of the n atoms and residues most rigid from mol, along the traj trajectory
// (see align/lovo.go:190-197 for the method and its references: 10.1371/journal.pone.0119264, 10.1186/1471-2105-8-306).
func LOVO(mol chem.Atomer, ref *v3.Matrix, trajname string, opt ...*Options) (*LOVOReturn, error) {
	var o *Options
	if len(opt) == 0 {
		o = DefaultOptions()
	} else {
		o = opt[0]
	}
	fullindexes := res2atoms(mol, o.atomNames, o.chains, nil)
	printtraj := o.writeTraj
	o.writeTraj = ""
	indexesold := make([]int, len(fullindexes))
	copy(indexesold, fullindexes)
	var RMSD *rMSD
	n := len(fullindexes)
	var itercount, disagreement int
	var prevlim float64
	// the file is read once; the frames stay on the device for every iteration
	load := func() (*deviceFrames, error) {
		traj, err := opentraj(trajname)
		if err != nil {
			return nil, err
		}
		defer traj.Close()
		return loadSelected(mol.Len(), traj, o)
	}
	d, err := load()
	if err != nil {
		return nil, err
	}
	defer func() { d.Close() }()
	// One device window and no aligned trajectory to write: the whole iteration runs in the library, the bookkeeping below on
	// the device (gcb_lovo_resident).  It declines what it does not cover; the loop below gives the same answers bit for bit.
	if printtraj == "" && len(d.windows) == 1 {
		res, taken, err := d.windows[0].LOVOResident(0, d.counts[0], ref, fullindexes, d.total, o.lessThanRMSD, o.minimumN, b200MaxIterations, nil)
		if err != nil {
			return nil, err
		}
		if taken {
			RMSD = newRMSD(res.IDs, res.Vals)
			RMSD.SortBy("molid")
			natoms := res.IndexesOld[0:res.N]
			return &LOVOReturn{N: res.N, Natoms: natoms, Nmols: iDs2MolIDandChains(mol, natoms), RMSD: RMSD.rMSDs, FramesRead: d.total,
				Iterations: res.Iterations}, nil
		}
	}
	for {
		if disagreement == 1 && printtraj != "" {
			o.writeTraj = printtraj
		}
		wb := o.writeTraj != ""
		rmsd, err := d.meanDeviations(ref, indexesold[0:n], wb)
		if err != nil {
			return nil, err
		}
		if wb {
			// this pass superposed the resident frames in place: write them, then start the next pass from the source again
			d.writeAligned(o.writeTraj)
			d.Close()
			if d, err = load(); err != nil {
				return nil, err
			}
		}
		itercount++
		RMSD = newRMSD(fullindexes, rmsdFilter(rmsd, fullindexes))
		RMSD.SortBy("rmsd")
		var newlim float64
		if o.lessThanRMSD > 0 {
			n, newlim = mobileLimit(o.lessThanRMSD, o.minimumN, RMSD)
		}
		indexes := RMSD.MolIDsCopy()
		if sameElementsInt(indexes[0:n], indexesold[0:n]) && (newlim == o.lessThanRMSD || approxEq(newlim, prevlim, 0.01)) {
			break // converged
		}
		prevlim = newlim
		disagreement = disagreementInt(indexes[0:n], indexesold[0:n])
		indexesold = indexes
	}
	RMSD.SortBy("molid")
	r := &LOVOReturn{N: n, Natoms: indexesold[0:n], Nmols: iDs2MolIDandChains(mol, indexesold[0:n]), RMSD: RMSD.rMSDs, FramesRead: d.total, Iterations: itercount}
	return r, nil
}
