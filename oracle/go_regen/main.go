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
// chem.RotatorTranslatorToSuper, chem.Super, chem.RMSD, chem.CenterOfMass, chem.MomentTensor, align.RMSDTraj (with Begin /
// Skip), align.LOVO (with and without Options.TrajName, the name of the aligned trajectory it writes), dcd.NewWriter / dcd.New and solv.FrameUMolCRDF -- one run pins every file
// of tests/golden, the reflection / planar / near-degenerate cases (super_edge_*) included.
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
//	<case>.moments.go.bin : per frame of a superposition case, float64 center[3] (chem.CenterOfMass, unit masses) and
//	                moment[9] (chem.MomentTensor), frame after frame
//	<case>.lovo.in.bin may be followed by int64 ncombos and int64 (begin, skip)[ncombos]:
//	<case>.rmsdtraj.go.bin : per combo, int64 framesRead and float64 rmsd[natoms] of align.RMSDTraj(fit on every second atom)
//	                with Options.Begin / Skip set (align/lovo.go:306-326)
//	<case>.lovo.aligned.dcd : the trajectory align.LOVO writes with Options.TrajName (lovo_options.go:137-142; lovo.go:209-211, 226-228, 296-305), byte
//	                for byte as dcd.DCDWObj produces it (only when the run reaches a disagreement of one atom)
//	<case>.rdf.in.bin : int64 natoms, nframes, nref; float64 step, end; int64 molid[natoms], chain[natoms], flags[natoms]
//	                (bit 0: residue name asked for, bit 1: water), refidx[nref]; float64 frames[nframes*natoms*3]
//	<case>.rdf.go.bin : float64 cdf[nframes][int(end/step)] of solv.FrameUMolCRDF without COM, then the same with Options.COM
package main

import (
	"encoding/binary"
	"fmt"
	"os"
	"path/filepath"
	"strings"

	chem "github.com/rmera/gochem"
	"github.com/rmera/gochem/align"
	"github.com/rmera/gochem/solv"
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

	// chem.CenterOfMass (unit masses) and chem.MomentTensor of every original frame (geometric.go:609-631, 705-733)
	mo, err := os.Create(strings.TrimSuffix(in, ".in.bin") + ".moments.go.bin")
	must(err)
	defer mo.Close()
	for fr := int64(0); fr < nframes; fr++ {
		x := matrix(clone(frames[fr*natoms*3 : (fr+1)*natoms*3]))
		c, err := chem.CenterOfMass(x)
		must(err)
		m, err := chem.MomentTensor(x)
		must(err)
		must(binary.Write(mo, binary.LittleEndian, c.RawSlice()))
		must(binary.Write(mo, binary.LittleEndian, m.RawSlice()))
	}
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

	// the same run with Options.TrajName set: the aligned trajectory is kept as written
	aligned := strings.TrimSuffix(in, ".lovo.in.bin") + ".lovo.aligned.dcd"
	optw := align.DefaultOptions()
	optw.Cpus(int(cpus))
	optw.Begin(0)
	optw.Skip(0)
	optw.LessThanRMSD(1.0)
	optw.MinimumN(10)
	optw.TrajName(aligned)
	_, err = align.LOVO(caTopology(natoms), matrix(clone(ref)), tmp, optw)
	must(err)

	// align.RMSDTraj with Begin / Skip on the DCD file, fit on every second atom
	var ncombos int64
	if binary.Read(f, binary.LittleEndian, &ncombos) != nil || ncombos <= 0 {
		return
	}
	combos := make([]int64, 2*ncombos)
	must(binary.Read(f, binary.LittleEndian, combos))
	rt, err := os.Create(strings.TrimSuffix(in, ".lovo.in.bin") + ".rmsdtraj.go.bin")
	must(err)
	defer rt.Close()
	sel := make([]int, 0, natoms/2+1)
	for i := 0; i < int(natoms); i += 2 {
		sel = append(sel, i)
	}
	for k := int64(0); k < ncombos; k++ {
		tr, err := dcd.New(tmp)
		must(err)
		o := align.DefaultOptions()
		o.Cpus(int(cpus))
		o.Begin(int(combos[2*k]))
		o.Skip(int(combos[2*k+1]))
		r, read, err := align.RMSDTraj(caTopology(natoms), matrix(clone(ref)), tr, sel, o)
		must(err)
		must(binary.Write(rt, binary.LittleEndian, int64(read)))
		must(binary.Write(rt, binary.LittleEndian, r))
	}
}

// solv.FrameUMolCRDF (solv/solvation.go:380-410) for every frame of a small solvated system, without and with Options.COM
func runRdfCase(in string) {
	f, err := os.Open(in)
	must(err)
	defer f.Close()
	var hdr [3]int64
	must(binary.Read(f, binary.LittleEndian, &hdr))
	natoms, nframes, nref := hdr[0], hdr[1], hdr[2]
	var se [2]float64
	must(binary.Read(f, binary.LittleEndian, &se))
	topo := make([]int64, 3*natoms)
	must(binary.Read(f, binary.LittleEndian, topo))
	ref64 := make([]int64, nref)
	must(binary.Read(f, binary.LittleEndian, ref64))
	frames := readF64(f, nframes*natoms*3)
	atoms := make([]*chem.Atom, natoms)
	for i := range atoms {
		a := new(chem.Atom)
		a.Name, a.Symbol, a.ID = "X", "C", i+1
		a.MolID = int(topo[i])
		a.Chain = fmt.Sprintf("%c", 'A'+rune(topo[natoms+int64(i)]))
		switch topo[2*natoms+int64(i)] {
		case 3:
			a.MolName = "SOL"
		case 1:
			a.MolName = "NA"
		default:
			a.MolName = "ALA"
		}
		atoms[i] = a
	}
	top := chem.NewTopology(0, 1, atoms)
	refidx := make([]int, nref)
	for i, v := range ref64 {
		refidx[i] = int(v)
	}
	g, err := os.Create(strings.TrimSuffix(in, ".in.bin") + ".go.bin")
	must(err)
	defer g.Close()
	for _, com := range []bool{false, true} {
		o := solv.DefaultOptions()
		o.Step(se[0])
		o.End(se[1])
		o.COM(com)
		for fr := int64(0); fr < nframes; fr++ {
			x := matrix(clone(frames[fr*natoms*3 : (fr+1)*natoms*3]))
			must(binary.Write(g, binary.LittleEndian, solv.FrameUMolCRDF(x, top, refidx, []string{"SOL", "NA"}, o)))
		}
	}
	fmt.Printf("%s: %d frames x %d atoms, %d solute atoms\n", filepath.Base(in), nframes, natoms, nref)
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
		} else if strings.HasSuffix(in, ".rdf.in.bin") {
			runRdfCase(in)
		} else {
			runCase(in)
		}
	}
}
