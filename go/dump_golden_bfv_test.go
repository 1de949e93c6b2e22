// dump_golden_bfv_test.go -- the mkbfv twin of dump_golden_test.go: drop it into mk-lattigo/mkbfv/ (package mkbfv) and run
//
//     MKHE_DUMP_DIR=/tmp/mkhe_dump_bfv go test ./mkbfv -run TestDumpGoldenBFV -args -n=4
//
// It writes the parameters (Q, QMul, P, T), lattigo's NTT tables of all three bases, CRS[-1], every party's relinearisation key
// (b1, d1, v | b2, d2), two input ciphertexts and the reference's own MulRelinNew output (mkbfv/evaluator.go:84-150: ModUpQtoR /
// Rescale, DecomposeBFV, MulAndRelinBFVHoisted, Quantize) as flat little-endian uint64 files plus meta.json ("scheme": "bfv"),
// in the layouts of include/mkhe.h.  `python tools/replay_dump.py /tmp/mkhe_dump_bfv` replays them through libmkhe_b200.so
// (mkhe_bfv_mul_relin) and through the CPU oracle and compares every output limb.
//
// NOT COMPILED OR RUN in this repository's environment (no Go toolchain); written against mkbfv/mkbfv_test.go:193-280
// (genTestParams, newTestVectors), mkbfv/keys.go:6-9 and mkbfv/evaluator.go:84-150.
package mkbfv

import (
	"encoding/binary"
	"encoding/json"
	"os"
	"path/filepath"
	"sort"
	"strconv"
	"testing"

	"github.com/ldsec/lattigo/v2/ring"
	"mk-lattigo/mkrlwe"
)

func bfvDumpU64(t *testing.T, dir, name string, rows ...[]uint64) {
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

// swk = [beta][nQ+nP][N]: digit-major, Q limbs then P limbs (mkrlwe/keys.go:23-25)
func bfvDumpSwk(t *testing.T, dir, name string, swk *mkrlwe.SwitchingKey) {
	rows := [][]uint64{}
	for d := range swk.Value {
		rows = append(rows, swk.Value[d].Q.Coeffs...)
		rows = append(rows, swk.Value[d].P.Coeffs...)
	}
	bfvDumpU64(t, dir, name, rows...)
}

func bfvDumpCt(t *testing.T, dir, prefix string, ct *Ciphertext, ids []string) {
	bfvDumpU64(t, dir, prefix+"_c0.bin", ct.Value["0"].Coeffs...)
	for i, id := range ids {
		bfvDumpU64(t, dir, prefix+"_p"+strconv.Itoa(i)+".bin", ct.Value[id].Coeffs...)
	}
}

func bfvDumpTables(t *testing.T, dir string, first int, r *ring.Ring) []uint64 {
	for i := range r.Modulus {
		bfvDumpU64(t, dir, "psi_"+strconv.Itoa(first+i)+".bin", r.NttPsi[i])
		bfvDumpU64(t, dir, "psiinv_"+strconv.Itoa(first+i)+".bin", r.NttPsiInv[i])
	}
	return r.NttNInv
}

func TestDumpGoldenBFV(t *testing.T) {
	dir := os.Getenv("MKHE_DUMP_DIR")
	if dir == "" {
		t.Skip("MKHE_DUMP_DIR not set")
	}
	if err := os.MkdirAll(dir, 0o755); err != nil {
		t.Fatal(err)
	}
	params := NewParametersFromLiteral(PN15QP880) // mkbfv_test.go:28-75: BASELINE config 3

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

	// inputs as in mkbfv_bench_test.go:39-64: sums of fresh encryptions of every party
	var ct0, ct1 *Ciphertext
	T := int64(params.T())
	for i, id := range ids {
		_, a := newTestVectors(tc, id, -T/4, T/4)
		_, b := newTestVectors(tc, id, -T/4, T/4)
		if i == 0 {
			ct0, ct1 = a, b
		} else {
			ct0, ct1 = tc.evaluator.AddNew(ct0, a), tc.evaluator.AddNew(ct1, b)
		}
	}

	ringQ, ringP, ringQMul := params.RingQ(), params.RingP(), params.RingQMul()
	nQ, nP := len(ringQ.Modulus), len(ringP.Modulus)
	ninv := append([]uint64{}, bfvDumpTables(t, dir, 0, ringQ)...)
	ninv = append(ninv, bfvDumpTables(t, dir, nQ, ringP)...)
	ninv = append(ninv, bfvDumpTables(t, dir, nQ+nP, ringQMul)...)
	bfvDumpU64(t, dir, "ninv.bin", ninv)
	bfvDumpSwk(t, dir, "crs_u.bin", params.CRS[-1])
	for i, id := range ids {
		rlk := tc.rlkSet.GetRelinearizationKey(id) // Value[0] = (b1, d1, v), Value[1] = (b2, d2, -)
		n := strconv.Itoa(i)
		bfvDumpSwk(t, dir, "rlk_"+n+"_b1.bin", rlk.Value[0].Value[0])
		bfvDumpSwk(t, dir, "rlk_"+n+"_d1.bin", rlk.Value[0].Value[1])
		bfvDumpSwk(t, dir, "rlk_"+n+"_v.bin", rlk.Value[0].Value[2])
		bfvDumpSwk(t, dir, "rlk_"+n+"_b2.bin", rlk.Value[1].Value[0])
		bfvDumpSwk(t, dir, "rlk_"+n+"_d2.bin", rlk.Value[1].Value[1])
	}
	bfvDumpCt(t, dir, "ct0", ct0, ids)
	bfvDumpCt(t, dir, "ct1", ct1, ids)

	mul := tc.evaluator.MulRelinNew(ct0, ct1, tc.rlkSet)
	bfvDumpCt(t, dir, "mul", mul, ids)

	meta := map[string]interface{}{
		"format": "mkhe-dump-1", "scheme": "bfv", "params": "BFV_PN15QP880", "logN": params.LogN(),
		"Q": ringQ.Modulus, "P": ringP.Modulus, "QMul": ringQMul.Modulus, "T": params.T(), "gamma": params.Gamma(),
		"ids": ids, "level": ct0.Level(), "source": "reference (Go, lattigo v2.3.0)",
	}
	js, _ := json.MarshalIndent(meta, "", " ")
	if err := os.WriteFile(filepath.Join(dir, "meta.json"), js, 0o644); err != nil {
		t.Fatal(err)
	}
}
