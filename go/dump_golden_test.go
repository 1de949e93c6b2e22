// dump_golden_test.go -- drop this file into mk-lattigo/mkckks/ (package mkckks) on a machine with Go and the lattigo v2.3.0
// module cache and run
//
//     MKHE_DUMP_DIR=/tmp/mkhe_dump go test ./mkckks -run TestDumpGolden -args -n=4
//
// It writes the parameters, lattigo's NTT tables, the CRS, every party's relinearisation and rotation key, two input
// ciphertexts and the reference's own outputs of MulRelinNew and RotateHoistedNew as flat little-endian uint64 files plus a
// meta.json, in exactly the layouts of include/mkhe.h.  `python tools/replay_dump.py /tmp/mkhe_dump` then replays the inputs
// through libmkhe_b200.so (and through the CPU oracle) and compares every output limb bit for bit: this is the out-of-band pin of
// "bit-exact with the Go evaluator" that cannot be produced inside the build container (no Go toolchain, DESIGN.md section 5).
//
// NOT COMPILED OR RUN in this repository's environment; written against the reference sources at mkckks/mkckks_test.go:135-179
// (genTestParams), mkckks/evaluator.go:416-443,543-617 and mkrlwe/keys.go.
package mkckks

import (
	"encoding/binary"
	"encoding/json"
	"os"
	"path/filepath"
	"sort"
	"strconv"
	"testing"

	"github.com/ldsec/lattigo/v2/ckks"
	"github.com/ldsec/lattigo/v2/ring"
	"mk-lattigo/mkrlwe"
)

func dumpU64(t *testing.T, dir, name string, rows ...[]uint64) {
	f, err := os.Create(filepath.Join(dir, name))
	if err != nil {
		t.Fatal(err)
	}
	defer f.Close()
	for _, r := range rows {
		if err := binary.Write(f, binary.LittleEndian, r); err != nil {
			t.Fatal(err)
		}
	}
}

// poly = [level+1][N]
func dumpPoly(t *testing.T, dir, name string, p *ring.Poly) { dumpU64(t, dir, name, p.Coeffs...) }

// swk = [beta][nQ+nP][N]: digit-major, Q limbs then P limbs (mkrlwe/keys.go:23-25)
func dumpSwk(t *testing.T, dir, name string, swk *mkrlwe.SwitchingKey) {
	rows := [][]uint64{}
	for d := range swk.Value {
		rows = append(rows, swk.Value[d].Q.Coeffs...)
		rows = append(rows, swk.Value[d].P.Coeffs...)
	}
	dumpU64(t, dir, name, rows...)
}

func dumpCt(t *testing.T, dir, prefix string, ct *Ciphertext, ids []string) {
	dumpPoly(t, dir, prefix+"_c0.bin", ct.Value["0"]) // component "0"; party t of the sorted id list -> _p<t>
	for i, id := range ids {
		dumpPoly(t, dir, prefix+"_p"+strconv.Itoa(i)+".bin", ct.Value[id])
	}
}

func TestDumpGolden(t *testing.T) {
	dir := os.Getenv("MKHE_DUMP_DIR")
	if dir == "" {
		t.Skip("MKHE_DUMP_DIR not set")
	}
	if err := os.MkdirAll(dir, 0o755); err != nil {
		t.Fatal(err)
	}
	ckksParams, err := ckks.NewParametersFromLiteral(PN15QP880)
	if err != nil {
		t.Fatal(err)
	}
	params := NewParameters(ckksParams)
	const rot = 2
	params.AddCRS(rot)

	ids := make([]string, *maxUsers)
	idset := mkrlwe.NewIDSet()
	for i := range ids {
		ids[i] = "user" + strconv.Itoa(i)
		idset.Add(ids[i])
	}
	sort.Strings(ids) // the C ABI takes parties in sorted order; index t in the dump = ids[t]

	tc, err := genTestParams(params, idset)
	if err != nil {
		t.Fatal(err)
	}
	for _, id := range ids {
		tc.rtkSet.AddRotationKey(tc.kgen.GenRotationKey(rot, tc.skSet.GetSecretKey(id)))
	}

	// inputs as in mkckks_benchmark_test.go:63-76: ct0 = sum of fresh encryptions, ct1 = another sum
	var ct0, ct1 *Ciphertext
	for i, id := range ids {
		_, a := newTestVectors(tc, id, complex(-1, -1), complex(1, 1))
		_, b := newTestVectors(tc, id, complex(-1, -1), complex(1, 1))
		if i == 0 {
			ct0, ct1 = a, b
		} else {
			ct0, ct1 = tc.evaluator.AddNew(ct0, a), tc.evaluator.AddNew(ct1, b)
		}
	}
	level := ct0.Level()

	ringQ, ringP := params.RingQ(), params.RingP()
	for i := range ringQ.Modulus {
		dumpU64(t, dir, "psi_"+strconv.Itoa(i)+".bin", ringQ.NttPsi[i])
		dumpU64(t, dir, "psiinv_"+strconv.Itoa(i)+".bin", ringQ.NttPsiInv[i])
	}
	for j := range ringP.Modulus {
		m := len(ringQ.Modulus) + j
		dumpU64(t, dir, "psi_"+strconv.Itoa(m)+".bin", ringP.NttPsi[j])
		dumpU64(t, dir, "psiinv_"+strconv.Itoa(m)+".bin", ringP.NttPsiInv[j])
	}
	dumpU64(t, dir, "ninv.bin", append(append([]uint64{}, ringQ.NttNInv...), ringP.NttNInv...))
	dumpSwk(t, dir, "crs_u.bin", params.CRS[-1])
	dumpSwk(t, dir, "crs_rot.bin", params.CRS[rot])
	for i, id := range ids {
		rlk := tc.rlkSet.GetRelinearizationKey(id)
		dumpSwk(t, dir, "rlk_"+strconv.Itoa(i)+"_b.bin", rlk.Value[0])
		dumpSwk(t, dir, "rlk_"+strconv.Itoa(i)+"_d.bin", rlk.Value[1])
		dumpSwk(t, dir, "rlk_"+strconv.Itoa(i)+"_v.bin", rlk.Value[2])
		dumpSwk(t, dir, "rk_"+strconv.Itoa(i)+".bin", tc.rtkSet.GetRotationKey(id, uint(rot)).Value)
	}
	dumpCt(t, dir, "ct0", ct0, ids)
	dumpCt(t, dir, "ct1", ct1, ids)

	// the reference's outputs.  MulRelinNew first COPIES nothing: ct0 / ct1 are read-only except that Rescale adds (q_l-1)/2 to
	// the last limb of its (freshly allocated) input, which is ctOut itself (mkckks/evaluator.go:385-390).
	mul := tc.evaluator.MulRelinNew(ct0, ct1, tc.rlkSet)
	dumpCt(t, dir, "mul", mul, ids)
	hoisted := tc.evaluator.HoistedForm(ct0)
	rotOut := tc.evaluator.RotateHoistedNew(ct0, rot, hoisted, tc.rtkSet)
	dumpCt(t, dir, "rot", rotOut, ids)

	meta := map[string]interface{}{
		"format": "mkhe-dump-1", "params": "PN15QP880", "logN": params.LogN(), "Q": ringQ.Modulus, "P": ringP.Modulus,
		"gamma": params.Gamma(), "scale": params.Scale(), "ids": ids, "level": level, "rot": rot,
		"mul_level": mul.Level(), "mul_scale": mul.Scale, "source": "reference (Go, lattigo v2.3.0)",
	}
	js, _ := json.MarshalIndent(meta, "", " ")
	if err := os.WriteFile(filepath.Join(dir, "meta.json"), js, 0o644); err != nil {
		t.Fatal(err)
	}
}
