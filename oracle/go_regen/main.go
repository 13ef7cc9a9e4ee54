// go_regen: pins the CPU oracle of this repository against the real goChem.
//
// This image has no Go toolchain, so the oracle (oracle/gochem_oracle.c, oracle/oracle_np.py) is "parity unpinned":
// nothing produced by the reference itself backs it.  Anyone with Go >= 1.22 and network access to the module proxy
// can close that gap:
//
//	python oracle/go_regen/export_inputs.py          # tests/golden/*.npz -> tests/golden/go_reference/<case>.in.bin
//	cd oracle/go_regen && go mod init go_regen && go get github.com/rmera/gochem@latest && go mod tidy
//	go run . ../../tests/golden/go_reference         # writes <case>.go.bin beside every <case>.in.bin
//	python -m pytest tests/test_oracle.py -k go_reference
//
// tests/test_oracle.py::test_go_reference_outputs and ::test_go_reference_lovo_output then compare the oracle with these
// files (they are skipped while the files are absent).  The program calls only the reference's public API:
// chem.RotatorTranslatorToSuper, chem.Super, chem.RMSD, align.RMSDTraj, align.LOVO and dcd.NewWriter.
//
// File formats (little endian).
//
//	<case>.in.bin : int64 natoms, nframes, nidx, cpus; float64 ref[natoms*3]; float64 frames[nframes*natoms*3];
//	                int64 idx[nidx]   (nidx = 0: all atoms)
//	<case>.go.bin : float64 rmsd[nframes]; rot[nframes*9]; trans1[nframes*3]; trans2[nframes*3];
//	                moved[nframes*natoms*3] (the frames after chem.Super); int64 framesRead;
//	                float64 rmsdtraj[natoms] (align.RMSDTraj over the ORIGINAL frames, chunk = cpus)
//	<case>.lovo.in.bin : the same header and ref / frames blocks (nidx = 0): the input of align.LOVO -- the frames are written
//	                to a DCD file first (dcd.NewWriter), every atom is a CA of chain A with MolID i+1, lessThanRMSD 1.0, minimumN 10
//	<case>.lovo.go.bin : int64 N, framesRead, iterations, nNatoms, nRMSD; int64 Natoms[nNatoms]; float64 RMSD[nRMSD]
package main

import (
	"encoding/binary"
	"fmt"
	"os"
	"path/filepath"
	"strings"

	chem "github.com/rmera/gochem"
	"github.com/rmera/gochem/align"
	"github.com/rmera/gochem/traj/dcd"
	v3 "github.com/rmera/gochem/v3"
)

func must(err error) {
	if err != nil {
		fmt.Fprintln(os.Stderr, err)
		os.Exit(1)
	}
}

func readF64(f *os.File, n int64) []float64 {
	out := make([]float64, n)
	must(binary.Read(f, binary.LittleEndian, out))
	return out
}

func matrix(data []float64) *v3.Matrix {
	m, err := v3.NewMatrix(data)
	must(err)
	return m
}

func clone(data []float64) []float64 {
	c := make([]float64, len(data))
	copy(c, data)
	return c
}

func runCase(in string) {
	f, err := os.Open(in)
	must(err)
	defer f.Close()
	var hdr [4]int64
	must(binary.Read(f, binary.LittleEndian, &hdr))
	natoms, nframes, nidx, cpus := hdr[0], hdr[1], hdr[2], hdr[3]
	ref := readF64(f, natoms*3)
	frames := readF64(f, nframes*natoms*3)
	idx64 := make([]int64, nidx)
	must(binary.Read(f, binary.LittleEndian, idx64))
	idx := make([]int, nidx)
	for i, v := range idx64 {
		idx[i] = int(v)
	}
	templa := matrix(clone(ref))

	rmsd := make([]float64, nframes)
	rot := make([]float64, 0, nframes*9)
	trans1 := make([]float64, 0, nframes*3)
	trans2 := make([]float64, 0, nframes*3)
	moved := make([]float64, 0, nframes*natoms*3)
	coords := make([]*v3.Matrix, nframes)
	for fr := int64(0); fr < nframes; fr++ {
		raw := frames[fr*natoms*3 : (fr+1)*natoms*3]
		coords[fr] = matrix(clone(raw)) // untouched copy for RMSDTraj below
		test := matrix(clone(raw))
		// rotation and translations of the fit subset, exactly as chem.Super obtains them (handy.go:175-214)
		ctest, ctempla := test, templa
		if nidx > 0 {
			ctest = v3.Zeros(len(idx))
			ctempla = v3.Zeros(len(idx))
			ctest.SomeVecs(test, idx)
			ctempla.SomeVecs(templa, idx)
		}
		_, r, t1, t2, err := chem.RotatorTranslatorToSuper(ctest, ctempla)
		must(err)
		rot = append(rot, r.RawSlice()...)
		trans1 = append(trans1, t1.RawSlice()...)
		trans2 = append(trans2, t2.RawSlice()...)
		var sup *v3.Matrix
		if nidx > 0 {
			sup, err = chem.Super(test, templa, idx, idx)
			must(err)
			rmsd[fr], err = chem.RMSD(sup, templa, idx, idx)
		} else {
			sup, err = chem.Super(test, templa)
			must(err)
			rmsd[fr], err = chem.RMSD(sup, templa)
		}
		must(err)
		moved = append(moved, sup.RawSlice()...)
	}

	// align.RMSDTraj over the original frames, served by a chem.Molecule (chem.go:729-755)
	atoms := make([]*chem.Atom, natoms)
	for i := range atoms {
		a := new(chem.Atom)
		a.Name = "CA"
		a.Symbol = "C"
		a.ID = i + 1
		a.MolID = i + 1
		a.MolName = "GLY"
		a.Chain = "A"
		atoms[i] = a
	}
	top := chem.NewTopology(0, 1, atoms)
	mol, err := chem.NewMolecule(coords, top, nil)
	must(err)
	opt := align.DefaultOptions()
	opt.Cpus(int(cpus))
	opt.Begin(0)
	opt.Skip(0)
	sel := idx
	if nidx == 0 {
		sel = make([]int, natoms)
		for i := range sel {
			sel[i] = i
		}
	}
	rt, read, err := align.RMSDTraj(top, templa, mol, sel, opt)
	must(err)

	out := strings.TrimSuffix(in, ".in.bin") + ".go.bin"
	g, err := os.Create(out)
	must(err)
	defer g.Close()
	for _, block := range [][]float64{rmsd, rot, trans1, trans2, moved} {
		must(binary.Write(g, binary.LittleEndian, block))
	}
	must(binary.Write(g, binary.LittleEndian, int64(read)))
	must(binary.Write(g, binary.LittleEndian, rt))
	fmt.Printf("%s: %d frames x %d atoms, fit on %d, RMSD[0] = %.12g, RMSDTraj read %d frames\n", filepath.Base(out), nframes, natoms, len(sel), rmsd[0], read)
}

func caTopology(natoms int64) *chem.Topology {
	atoms := make([]*chem.Atom, natoms)
	for i := range atoms {
		a := new(chem.Atom)
		a.Name = "CA"
		a.Symbol = "C"
		a.ID = i + 1
		a.MolID = i + 1
		a.MolName = "GLY"
		a.Chain = "A"
		atoms[i] = a
	}
	return chem.NewTopology(0, 1, atoms)
}

// align.LOVO on a DCD file written from the frames of the case (align/lovo.go:198-258)
func runLovoCase(in string) {
	f, err := os.Open(in)
	must(err)
	defer f.Close()
	var hdr [4]int64
	must(binary.Read(f, binary.LittleEndian, &hdr))
	natoms, nframes, cpus := hdr[0], hdr[1], hdr[3]
	ref := readF64(f, natoms*3)
	frames := readF64(f, nframes*natoms*3)
	tmp := strings.TrimSuffix(in, ".lovo.in.bin") + ".lovo.tmp.dcd"
	w, err := dcd.NewWriter(tmp, int(natoms))
	must(err)
	for fr := int64(0); fr < nframes; fr++ {
		must(w.WNext(matrix(clone(frames[fr*natoms*3 : (fr+1)*natoms*3]))))
	}
	w.Close()
	defer os.Remove(tmp)
	opt := align.DefaultOptions()
	opt.Cpus(int(cpus))
	opt.Begin(0)
	opt.Skip(0)
	opt.LessThanRMSD(1.0)
	opt.MinimumN(10)
	ret, err := align.LOVO(caTopology(natoms), matrix(clone(ref)), tmp, opt)
	must(err)
	out := strings.TrimSuffix(in, ".in.bin") + ".go.bin"
	g, err := os.Create(out)
	must(err)
	defer g.Close()
	nat := make([]int64, len(ret.Natoms))
	for i, v := range ret.Natoms {
		nat[i] = int64(v)
	}
	must(binary.Write(g, binary.LittleEndian, []int64{int64(ret.N), int64(ret.FramesRead), int64(ret.Iterations), int64(len(nat)), int64(len(ret.RMSD))}))
	must(binary.Write(g, binary.LittleEndian, nat))
	must(binary.Write(g, binary.LittleEndian, ret.RMSD))
	fmt.Printf("%s: N = %d, %d frames read, %d iterations\n", filepath.Base(out), ret.N, ret.FramesRead, ret.Iterations)
}

func main() {
	dir := "../../tests/golden/go_reference"
	if len(os.Args) > 1 {
		dir = os.Args[1]
	}
	ins, err := filepath.Glob(filepath.Join(dir, "*.in.bin"))
	must(err)
	if len(ins) == 0 {
		must(fmt.Errorf("no *.in.bin under %s: run oracle/go_regen/export_inputs.py first", dir))
	}
	for _, in := range ins {
		if strings.HasSuffix(in, ".lovo.in.bin") {
			runLovoCase(in)
		} else {
			runCase(in)
		}
	}
}
