#!/bin/sh
# apply.sh -- put the b200 drop-in files into a checkout of github.com/rmera/gochem (the tree this repository calls
# /root/reference, go.mod: module github.com/rmera/gochem, gonum v0.15.1).
#
#   sh go/apply.sh /path/to/gochem /path/to/this/repo
#
# It (1) copies package v3gpu and the two *_b200.go files, (2) MOVES the CPU bodies they replace out of handy.go,
# geometric.go and align/lovo.go into *_cpu.go files tagged `!b200` -- by line range, verified against the function
# signatures first, so nothing of the reference is retyped here -- and (3) copies the C header next to the binding.
# `go build ./...` is then the unmodified reference; `go build -tags b200 ./...` binds libgochem_b200.so:
#
#   CGO_LDFLAGS="-L$REPO/gochem_b200/lib -Wl,-rpath,$REPO/gochem_b200/lib" go build -tags b200 ./...
#
# Users of the module switch with one line in their go.mod:   replace github.com/rmera/gochem => /path/to/gochem
set -eu
GO=${1:?path to the goChem checkout}
REPO=${2:?path to this repository}

expect() {  # file line pattern
    sed -n "$2p" "$1" | grep -q "$3" || { echo "apply.sh: $1:$2 does not hold '$3' -- different revision of goChem?" >&2; exit 1; }
}
# move lines [a,b] of file $1 into $2 (appending), then blank them out of $1 at the end (ranges are given bottom-up)
cut_range() { sed -n "$3,$4p" "$1" >> "$2"; sed -i "$3,$4d" "$1"; }

# ---- chem: handy.go (Super) and geometric.go (RotatorTranslatorToSuper, RMSD, MemRMSD, MemPerAtomRMSD) ----
expect "$GO/handy.go" 175 '^func Super(test, templa \*v3.Matrix, indexes \.\.\.\[\]int)'
expect "$GO/handy.go" 214 '^}'
expect "$GO/geometric.go" 127 '^func RotatorTranslatorToSuper(test, templa \*v3.Matrix)'
expect "$GO/geometric.go" 208 '^}'
expect "$GO/geometric.go" 232 '^func RMSD(test, templa \*v3.Matrix, indexes \.\.\.\[\]int)'
expect "$GO/geometric.go" 258 '^func MemRMSD(test, templa, tmp \*v3.Matrix, indexes \.\.\.\[\]int)'
expect "$GO/geometric.go" 294 '^func MemPerAtomRMSD(test, templa, ctest, ctempla, tmp \*v3.Matrix, indexes \.\.\.\[\]int)'
expect "$GO/geometric.go" 321 '^}'
CPU="$GO/super_cpu.go"
printf '//go:build !b200\n\npackage chem\n\nimport (\n\t"fmt"\n\t"math"\n\n\t"gonum.org/v1/gonum/mat"\n\n\tv3 "github.com/rmera/gochem/v3"\n)\n\n' > "$CPU"
# bottom-up inside each file so earlier line numbers stay valid; doc comments travel with their functions
cut_range "$GO/geometric.go" "$CPU" 290 321     # MemPerAtomRMSD
cut_range "$GO/geometric.go" "$CPU" 245 288     # MemRMSD
cut_range "$GO/geometric.go" "$CPU" 225 243     # RMSD
cut_range "$GO/geometric.go" "$CPU" 120 208     # RotatorTranslatorToSuper
cut_range "$GO/handy.go" "$CPU" 168 214         # Super
cp "$REPO/go/chem/super_b200.go" "$GO/super_b200.go"

# ---- align: lovo.go (LOVO, rmsdandcoords, concproc, RMSDTraj) ----
expect "$GO/align/lovo.go" 198 '^func LOVO(mol chem.Atomer, ref \*v3.Matrix, trajname string, opt \.\.\.\*Options)'
expect "$GO/align/lovo.go" 265 '^func concproc('
expect "$GO/align/lovo.go" 286 '^func RMSDTraj(mol chem.Atomer, ref \*v3.Matrix, traj chem.ConcTraj, indexes \[\]int, o \*Options)'
expect "$GO/align/lovo.go" 353 '^}'
ACPU="$GO/align/lovo_cpu.go"
printf '//go:build !b200\n\npackage align\n\nimport (\n\t"fmt"\n\t"log"\n\n\tchem "github.com/rmera/gochem"\n\t"github.com/rmera/gochem/traj/dcd"\n\tv3 "github.com/rmera/gochem/v3"\n)\n\n' > "$ACPU"
cut_range "$GO/align/lovo.go" "$ACPU" 190 353   # LOVO's doc comment .. end of RMSDTraj
cp "$REPO/go/align/lovo_b200.go" "$GO/align/lovo_b200.go"

# ---- the binding ----
mkdir -p "$GO/v3gpu"
cp "$REPO/go/v3gpu/v3gpu.go" "$GO/v3gpu/v3gpu.go"
cp "$REPO/include/gochem_b200.h" "$GO/v3gpu/gochem_b200.h"

# imports that only the moved bodies used are now unused in the originals: let the toolchain say which, then drop them
if command -v goimports >/dev/null 2>&1; then
    goimports -w "$GO/handy.go" "$GO/geometric.go" "$GO/align/lovo.go" "$CPU" "$ACPU"
else
    echo "apply.sh: run goimports (or gofmt + remove unused imports) on handy.go, geometric.go, align/lovo.go, super_cpu.go, align/lovo_cpu.go" >&2
fi
echo "applied: go vet ./... && go vet -tags b200 ./..."
