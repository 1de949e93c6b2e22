// Package mkgpu is the cgo binding of libmkhe_b200.so (include/mkhe.h) for mk-lattigo: device contexts, device-resident
// polynomials / switching keys and one Go method per C entry point of the hot path.  The method bodies of mkrlwe.KeySwitcher,
// mkckks.Evaluator and mkbfv.Evaluator call these (go/patches/*.patch); every exported type and signature of the reference
// stays as it is.
//
// NOT COMPILED in this repository's build container (no Go toolchain, lattigo v2.3.0 not vendored): written against
// include/mkhe.h, whose entry points the parity tests exercise through ctypes with the same argument conventions.
//
//	export CGO_CFLAGS="-I$REPO/include"
//	export CGO_LDFLAGS="-L$REPO/mkhe_kklss_b200 -lmkhe_b200 -Wl,-rpath,$REPO/mkhe_kklss_b200"
package mkgpu

/*
#include <stdlib.h>
#include <mkhe.h>
*/
import "C"

import (
	"errors"
	"runtime"
	"sort"
	"unsafe"

	"github.com/ldsec/lattigo/v2/ring"
	"github.com/ldsec/lattigo/v2/rlwe"
)

// Ctx is one device context (= one lane: a stream plus the scratch pools the reference keeps in KeySwitcher / HoistPool).
// Like the reference's KeySwitcher it is not goroutine-safe; a second evaluator over the same parameters is a Fork.
type Ctx struct {
	p      *C.mkhe_ctx
	nQ, nP int
	n      int
}

// NewCtx mirrors mkrlwe.NewParameters: moduli, gamma (alpha = #P / gamma digits), device ordinal.  lattigo's own NTT tables are
// handed over so that the psi convention is lattigo's by construction (the library's generator follows the same rule).
func NewCtx(params rlwe.Parameters, gamma, device int) *Ctx {
	q, p := params.Q(), params.P()
	var c *C.mkhe_ctx
	rc := C.mkhe_ctx_create(C.int(params.LogN()), (*C.uint64_t)(&q[0]), C.int(len(q)), (*C.uint64_t)(&p[0]), C.int(len(p)),
		C.int(gamma), C.int(device), &c)
	if rc != 0 {
		panic("mkgpu: mkhe_ctx_create failed (no CUDA device? more than 4 special primes?)")
	}
	ctx := &Ctx{p: c, nQ: len(q), nP: len(p), n: params.N()}
	for i := range q {
		ctx.setTables(i, params.RingQ(), i)
	}
	for j := range p {
		ctx.setTables(len(q)+j, params.RingP(), j)
	}
	runtime.SetFinalizer(ctx, func(x *Ctx) { C.mkhe_ctx_destroy(x.p) })
	return ctx
}

// SetBFV adds the auxiliary basis QMul and the plaintext modulus (mkbfv.NewParametersFromLiteral).
func (c *Ctx) SetBFV(ringQMul *ring.Ring, t uint64) {
	qm := ringQMul.Modulus
	c.check(C.mkhe_ctx_set_bfv(c.p, (*C.uint64_t)(&qm[0]), C.int(len(qm)), C.uint64_t(t)))
	for i := range qm {
		c.setTables(c.nQ+c.nP+i, ringQMul, i)
	}
}

func (c *Ctx) setTables(m int, r *ring.Ring, i int) {
	c.check(C.mkhe_ctx_set_ntt_tables(c.p, C.int(m), (*C.uint64_t)(&r.NttPsi[i][0]), (*C.uint64_t)(&r.NttPsiInv[i][0]), C.uint64_t(r.NttNInv[i])))
}

func (c *Ctx) check(rc C.int) {
	if rc != 0 {
		panic(errors.New(C.GoString(C.mkhe_last_error(c.p))))
	}
}

// Fork = a second NewEvaluator / NewKeySwitcher over the same Parameters (mkrlwe/keyswitch.go:33-47): same handles, own stream
// and pools.  Independent ops on the two lanes overlap on the device; uses of one object on both are ordered by the library.
func (c *Ctx) Fork() *Ctx {
	var f *C.mkhe_ctx
	c.check(C.mkhe_ctx_fork(c.p, &f))
	return &Ctx{p: f, nQ: c.nQ, nP: c.nP, n: c.n}
}
func (c *Ctx) Sync() { c.check(C.mkhe_sync(c.p)) }

// ---- device objects ------------------------------------------------------------------------------------------------------

// Poly mirrors one *ring.Poly; every Coeffs[i] is its own Go slice, hence per-limb transfers.
type Poly struct {
	h C.mkhe_poly
	c *Ctx
}

func (c *Ctx) NewPoly(capLimbs int) *Poly {
	d := &Poly{c: c}
	c.check(C.mkhe_poly_alloc(c.p, C.int(capLimbs), &d.h))
	runtime.SetFinalizer(d, func(x *Poly) { C.mkhe_poly_free(x.c.p, x.h) }) // stream-ordered: the block returns to the pool
	return d
}
func (c *Ctx) UploadPoly(p *ring.Poly, capLimbs int) *Poly {
	d := c.NewPoly(capLimbs)
	d.Upload(p)
	return d
}
func (d *Poly) Upload(p *ring.Poly) {
	for i := range p.Coeffs {
		d.c.check(C.mkhe_poly_upload_limb(d.c.p, d.h, C.int(i), (*C.uint64_t)(unsafe.Pointer(&p.Coeffs[i][0]))))
	}
	d.SetLevel(len(p.Coeffs) - 1)
}
func (d *Poly) Download(p *ring.Poly) {
	for i := range p.Coeffs {
		d.c.check(C.mkhe_poly_download_limb(d.c.p, d.h, C.int(i), (*C.uint64_t)(unsafe.Pointer(&p.Coeffs[i][0]))))
	}
}
func (d *Poly) SetLevel(level int) { d.c.check(C.mkhe_poly_set_nlimbs(d.c.p, d.h, C.int(level+1))) }

// UploadQP uploads one rlwe.PolyQP (a secret or public key) as a poly of nQ + nP limbs, Q limbs first.
func (c *Ctx) UploadQP(v rlwe.PolyQP) *Poly {
	d := c.NewPoly(c.nQ + c.nP)
	for i, l := range v.Q.Coeffs {
		c.check(C.mkhe_poly_upload_limb(c.p, d.h, C.int(i), (*C.uint64_t)(unsafe.Pointer(&l[0]))))
	}
	for i, l := range v.P.Coeffs {
		c.check(C.mkhe_poly_upload_limb(c.p, d.h, C.int(c.nQ+i), (*C.uint64_t)(unsafe.Pointer(&l[0]))))
	}
	return d
}

// Swk mirrors one mkrlwe.SwitchingKey (Value []rlwe.PolyQP): a key, a CRS entry or one hoisted form.
type Swk struct {
	h C.mkhe_swk
	c *Ctx
}

func (c *Ctx) NewSwk() *Swk {
	d := &Swk{c: c}
	c.check(C.mkhe_swk_alloc(c.p, &d.h))
	runtime.SetFinalizer(d, func(x *Swk) { C.mkhe_swk_free(x.c.p, x.h) })
	return d
}
func (c *Ctx) UploadSwk(v []rlwe.PolyQP) *Swk {
	d := c.NewSwk()
	for dig := range v {
		for i, l := range v[dig].Q.Coeffs {
			c.check(C.mkhe_swk_upload_limb(c.p, d.h, C.int(dig), 0, C.int(i), (*C.uint64_t)(unsafe.Pointer(&l[0]))))
		}
		for i, l := range v[dig].P.Coeffs {
			c.check(C.mkhe_swk_upload_limb(c.p, d.h, C.int(dig), 1, C.int(i), (*C.uint64_t)(unsafe.Pointer(&l[0]))))
		}
	}
	return d
}
func (d *Swk) Download(v []rlwe.PolyQP) {
	for dig := range v {
		for i, l := range v[dig].Q.Coeffs {
			d.c.check(C.mkhe_swk_download_limb(d.c.p, d.h, C.int(dig), 0, C.int(i), (*C.uint64_t)(unsafe.Pointer(&l[0]))))
		}
		for i, l := range v[dig].P.Coeffs {
			d.c.check(C.mkhe_swk_download_limb(d.c.p, d.h, C.int(dig), 1, C.int(i), (*C.uint64_t)(unsafe.Pointer(&l[0]))))
		}
	}
}

// HostAlloc returns page-locked memory owned by the library (C memory: the cgo pointer rule does not apply), the backing store
// for Coeffs of ciphertexts that stream through UploadAsync / DownloadAsync.
func (c *Ctx) HostAlloc(n int) []uint64 {
	var p unsafe.Pointer
	c.check(C.mkhe_host_alloc(c.p, C.size_t(8*n), &p))
	return unsafe.Slice((*uint64)(p), n)
}
func (d *Poly) UploadAsync(pinned []uint64, nlimbs int) {
	d.c.check(C.mkhe_poly_upload_async(d.c.p, d.h, (*C.uint64_t)(unsafe.Pointer(&pinned[0])), C.int(nlimbs)))
}
func (d *Poly) DownloadAsync(pinned []uint64, nlimbs int) {
	d.c.check(C.mkhe_poly_download_async(d.c.p, d.h, (*C.uint64_t)(unsafe.Pointer(&pinned[0])), C.int(nlimbs)))
}

// ---- handle lists ---------------------------------------------------------------------------------------------------------

// SortedIDs returns the party ids of a Go map in the order the C ABI takes them (sorted; index t in every handle list).
func SortedIDs(m map[string]*Poly) []string {
	ids := make([]string, 0, len(m))
	for id := range m {
		if id != "0" {
			ids = append(ids, id)
		}
	}
	sort.Strings(ids)
	return ids
}

// the C ABI identifies parties by small integers; the shim numbers the union of the ids of a call
func number(idsets ...[]string) (map[string]C.int, []string) {
	all := map[string]bool{}
	for _, s := range idsets {
		for _, id := range s {
			all[id] = true
		}
	}
	u := make([]string, 0, len(all))
	for id := range all {
		u = append(u, id)
	}
	sort.Strings(u)
	num := map[string]C.int{}
	for i, id := range u {
		num[id] = C.int(i)
	}
	return num, u
}
func ints(num map[string]C.int, ids []string) []C.int {
	o := make([]C.int, len(ids)+1)
	for i, id := range ids {
		o[i] = num[id]
	}
	return o
}
func ctHandles(ct map[string]*Poly, ids []string) []C.mkhe_poly {
	o := make([]C.mkhe_poly, 0, len(ids)+1)
	o = append(o, ct["0"].h)
	for _, id := range ids {
		o = append(o, ct[id].h)
	}
	return o
}
func swkHandles(m map[string]*Swk, ids []string) []C.mkhe_swk {
	if m == nil {
		return nil
	}
	o := make([]C.mkhe_swk, len(ids)+1)
	for i, id := range ids {
		o[i] = m[id].h
	}
	return o
}
func swkPtr(s []C.mkhe_swk) *C.mkhe_swk {
	if s == nil {
		return nil // Go nil hoisted form: the library decomposes the operand itself (keyswitch.go:122-230)
	}
	return &s[0]
}

// ---- mkrlwe.KeySwitcher ---------------------------------------------------------------------------------------------------

func (c *Ctx) Decompose(levelQ int, a *Poly, ad *Swk) { c.check(C.mkhe_decompose(c.p, C.int(levelQ), a.h, ad.h)) }
func (c *Ctx) ExternalProduct(levelQ int, a *Poly, bg *Swk, out *Poly) {
	c.check(C.mkhe_external_product(c.p, C.int(levelQ), a.h, bg.h, out.h))
}
func (c *Ctx) ExternalProductHoisted(levelQ int, ah *Swk, bg *Swk, out *Poly) {
	c.check(C.mkhe_external_product_hoisted(c.p, C.int(levelQ), ah.h, bg.h, out.h))
}

// MulAndRelinHoisted: op0 / op1 / out are Ciphertext.Value maps of device polys; h0 / h1 the hoisted forms or nil; b, d, v map a
// party id to the three switching keys of its RelinearizationKey; u = CRS[-1].
func (c *Ctx) MulAndRelinHoisted(level int, op0, op1 map[string]*Poly, h0, h1 map[string]*Swk, b, d, v map[string]*Swk, u *Swk, out map[string]*Poly) {
	ids0, ids1, idsO := SortedIDs(op0), SortedIDs(op1), SortedIDs(out)
	num, _ := number(ids0, ids1, idsO)
	i0, i1, iO := ints(num, ids0), ints(num, ids1), ints(num, idsO)
	o0, o1, oo := ctHandles(op0, ids0), ctHandles(op1, ids1), ctHandles(out, idsO)
	kb, kd, kv := swkHandles(b, ids1), swkHandles(d, ids0), swkHandles(v, ids0)
	c.check(C.mkhe_mul_relin_hoisted(c.p, C.int(level), C.int(len(ids0)), &i0[0], &o0[0], swkPtr(swkHandles(h0, ids0)),
		C.int(len(ids1)), &i1[0], &o1[0], swkPtr(swkHandles(h1, ids1)), &kb[0], &kd[0], &kv[0], u.h, C.int(len(idsO)), &iO[0], &oo[0]))
}

// CkksMulRelin = mkckks.Evaluator.MulRelinNew in one call: hoist both operands (pools in the context, like rlkSet.HoistPool),
// MulAndRelinHoisted, Rescale by nbRescales limbs.
func (c *Ctx) CkksMulRelin(level, nbRescales int, same bool, op0, op1 map[string]*Poly, b, d, v map[string]*Swk, u *Swk, out map[string]*Poly) {
	ids0, ids1, idsO := SortedIDs(op0), SortedIDs(op1), SortedIDs(out)
	num, _ := number(ids0, ids1, idsO)
	i0, i1, iO := ints(num, ids0), ints(num, ids1), ints(num, idsO)
	o0, o1, oo := ctHandles(op0, ids0), ctHandles(op1, ids1), ctHandles(out, idsO)
	kb, kd, kv := swkHandles(b, ids1), swkHandles(d, ids0), swkHandles(v, ids0)
	s := C.int(0)
	if same {
		s = 1
	}
	c.check(C.mkhe_ckks_mul_relin(c.p, C.int(level), C.int(nbRescales), s, C.int(len(ids0)), &i0[0], &o0[0], C.int(len(ids1)), &i1[0], &o1[0],
		&kb[0], &kd[0], &kv[0], u.h, C.int(len(idsO)), &iO[0], &oo[0]))
}

func (c *Ctx) RotateHoisted(level, rotidx int, in map[string]*Poly, hoisted, rk map[string]*Swk, a *Swk, out map[string]*Poly) {
	ids := SortedIDs(in)
	ci, co := ctHandles(in, ids), ctHandles(out, ids)
	hh, kk := swkHandles(hoisted, ids), swkHandles(rk, ids)
	c.check(C.mkhe_rotate_hoisted(c.p, C.int(level), C.int(rotidx), C.int(len(ids)), &ci[0], &hh[0], &kk[0], a.h, &co[0]))
}
func (c *Ctx) Rotate(level, rotidx int, in map[string]*Poly, rk map[string]*Swk, a *Swk, out map[string]*Poly) {
	ids := SortedIDs(in)
	ci, co, kk := ctHandles(in, ids), ctHandles(out, ids), swkHandles(rk, ids)
	c.check(C.mkhe_rotate(c.p, C.int(level), C.int(rotidx), C.int(len(ids)), &ci[0], &kk[0], a.h, &co[0]))
}
func (c *Ctx) Conjugate(level int, in map[string]*Poly, ck map[string]*Swk, a *Swk, out map[string]*Poly) {
	ids := SortedIDs(in)
	ci, co, kk := ctHandles(in, ids), ctHandles(out, ids), swkHandles(ck, ids)
	c.check(C.mkhe_conjugate(c.p, C.int(level), C.int(len(ids)), &ci[0], &kk[0], a.h, &co[0]))
}

// ---- mkckks.Evaluator: element-wise pieces --------------------------------------------------------------------------------

func (c *Ctx) Rescale(level, nb int, in, out *Poly) { c.check(C.mkhe_rescale(c.p, C.int(level), C.int(nb), in.h, out.h)) }
func (c *Ctx) Add(level int, a, b, out *Poly)       { c.check(C.mkhe_poly_add(c.p, C.int(level), a.h, b.h, out.h)) }
func (c *Ctx) Sub(level int, a, b, out *Poly)       { c.check(C.mkhe_poly_sub(c.p, C.int(level), a.h, b.h, out.h)) }
func (c *Ctx) Neg(level int, a, out *Poly)          { c.check(C.mkhe_poly_neg(c.p, C.int(level), a.h, out.h)) }
func (c *Ctx) CopyLvl(level int, dst, src *Poly)    { c.check(C.mkhe_poly_copy_lvl(c.p, C.int(level), dst.h, src.h)) }
func (c *Ctx) MultByConst(level int, in, out []*Poly, cReal, cImag, scale float64) {
	hi, ho := polyHandles(in), polyHandles(out)
	c.check(C.mkhe_ckks_mult_by_const(c.p, C.int(level), C.int(len(in)), &hi[0], &ho[0], C.double(cReal), C.double(cImag), C.double(scale)))
}
func (c *Ctx) MulPtxt(level int, pt *Poly, in, out []*Poly) {
	hi, ho := polyHandles(in), polyHandles(out)
	c.check(C.mkhe_ckks_mul_ptxt(c.p, C.int(level), pt.h, C.int(len(in)), &hi[0], &ho[0]))
}
func polyHandles(ps []*Poly) []C.mkhe_poly {
	o := make([]C.mkhe_poly, len(ps)+1)
	for i, p := range ps {
		o[i] = p.h
	}
	return o
}

// Decrypt (mkrlwe/decryptor.go:48-66): sk maps a party id to its SecretKey.Value uploaded with UploadQP.
func (c *Ctx) Decrypt(level int, ct map[string]*Poly, sk map[string]*Poly, pt *Poly) {
	ids := SortedIDs(ct)
	cc := ctHandles(ct, ids)
	ks := make([]C.mkhe_poly, len(ids)+1)
	for i, id := range ids {
		s, ok := sk[id]
		if !ok {
			panic("Cannot Decrypt: there is a missing secretkey")
		}
		ks[i] = s.h
	}
	c.check(C.mkhe_decrypt(c.p, C.int(level), C.int(len(ids)), &cc[0], &ks[0], pt.h))
}

// ---- mkbfv ----------------------------------------------------------------------------------------------------------------

func (c *Ctx) BfvModUpQtoR(q, r *Poly)  { c.check(C.mkhe_bfv_modup_q_to_r(c.p, q.h, r.h)) }
func (c *Ctx) BfvRescale(q, r *Poly)    { c.check(C.mkhe_bfv_rescale_q_to_r(c.p, q.h, r.h)) }
func (c *Ctx) BfvQuantize(r, q *Poly)   { c.check(C.mkhe_bfv_quantize(c.p, r.h, q.h)) }
func (c *Ctx) BfvDecompose(level int, r *Poly, ad1, ad2 *Swk) {
	c.check(C.mkhe_bfv_decompose(c.p, C.int(level), r.h, ad1.h, ad2.h))
}

// BfvMulRelin = mkbfv.Evaluator.MulRelinNew's device work (mkbfv/evaluator.go:84-150).
func (c *Ctx) BfvMulRelin(ct0, ct1 map[string]*Poly, b1, b2, d1, d2, v map[string]*Swk, u *Swk, out map[string]*Poly) {
	ids0, ids1, idsO := SortedIDs(ct0), SortedIDs(ct1), SortedIDs(out)
	num, _ := number(ids0, ids1, idsO)
	i0, i1, iO := ints(num, ids0), ints(num, ids1), ints(num, idsO)
	o0, o1, oo := ctHandles(ct0, ids0), ctHandles(ct1, ids1), ctHandles(out, idsO)
	kb1, kb2 := swkHandles(b1, ids1), swkHandles(b2, ids1)
	kd1, kd2, kv := swkHandles(d1, ids0), swkHandles(d2, ids0), swkHandles(v, ids0)
	c.check(C.mkhe_bfv_mul_relin(c.p, C.int(len(ids0)), &i0[0], &o0[0], C.int(len(ids1)), &i1[0], &o1[0],
		&kb1[0], &kb2[0], &kd1[0], &kd2[0], &kv[0], u.h, C.int(len(idsO)), &iO[0], &oo[0]))
}

// ---- key generation and encryption on the device (counter-based streams "mkhe-ctr-1", include/mkhe_prng.h) -----------------

// KeyGen holds (seed, next stream): every call consumes the number of streams include/mkhe.h documents.
type KeyGen struct {
	c            *Ctx
	Seed, Stream uint64
	beta         uint64
}

func (c *Ctx) NewKeyGen(seed, stream uint64, beta int) *KeyGen { return &KeyGen{c, seed, stream, uint64(beta)} }
func (g *KeyGen) take(n uint64) C.uint64_t                      { s := g.Stream; g.Stream += n; return C.uint64_t(s) }
func (c *Ctx) SampleCRS(seed, stream uint64, out *Swk)          { c.check(C.mkhe_sample_crs(c.p, C.uint64_t(seed), C.uint64_t(stream), out.h)) }
func (g *KeyGen) GenSecretKey(pZero float64) *Poly {
	sk := g.c.NewPoly(g.c.nQ + g.c.nP)
	g.c.check(C.mkhe_keygen_secret(g.c.p, C.uint64_t(g.Seed), g.take(1), C.double(pZero), sk.h))
	return sk
}
func (g *KeyGen) GenPublicKey(sk *Poly, crsA *Swk) (pk0, pk1 *Poly) {
	pk0, pk1 = g.c.NewPoly(g.c.nQ+g.c.nP), g.c.NewPoly(g.c.nQ+g.c.nP)
	g.c.check(C.mkhe_keygen_public(g.c.p, C.uint64_t(g.Seed), g.take(1), sk.h, crsA.h, pk0.h, pk1.h))
	return
}
func (g *KeyGen) GenSwitchingKey(sk *Poly, out *Swk) {
	g.c.check(C.mkhe_keygen_switching_key(g.c.p, C.uint64_t(g.Seed), g.take(g.beta), sk.h, out.h))
}
func (g *KeyGen) GenRelinearizationKey(sk, r *Poly, crsA, crsU *Swk) (b, d, v *Swk) {
	b, d, v = g.c.NewSwk(), g.c.NewSwk(), g.c.NewSwk()
	g.c.check(C.mkhe_keygen_relin(g.c.p, C.uint64_t(g.Seed), g.take(3*g.beta), sk.h, r.h, crsA.h, crsU.h, b.h, d.h, v.h))
	return
}
func (g *KeyGen) GenRotationKey(rotidx int, sk *Poly, crsRot *Swk) *Swk {
	rk := g.c.NewSwk()
	g.c.check(C.mkhe_keygen_rotation(g.c.p, C.uint64_t(g.Seed), g.take(g.beta), C.int(rotidx), sk.h, crsRot.h, rk.h))
	return rk
}
func (g *KeyGen) GenConjugationKey(sk *Poly, crsCj *Swk) *Swk {
	ck := g.c.NewSwk()
	g.c.check(C.mkhe_keygen_conjugation(g.c.p, C.uint64_t(g.Seed), g.take(g.beta), sk.h, crsCj.h, ck.h))
	return ck
}
func (g *KeyGen) GenBFVRelinearizationKey(sk, r *Poly, a1, a2, u *Swk) (b1, b2, d1, d2, v *Swk) {
	b1, b2, d1, d2, v = g.c.NewSwk(), g.c.NewSwk(), g.c.NewSwk(), g.c.NewSwk(), g.c.NewSwk()
	g.c.check(C.mkhe_keygen_bfv_relin(g.c.p, C.uint64_t(g.Seed), g.take(5*g.beta), sk.h, r.h, a1.h, a2.h, u.h, b1.h, b2.h, d1.h, d2.h, v.h))
	return
}

// Encrypt (mkrlwe/encryptor.go:55-118, coefficient-domain ciphertext); pt may be nil (encryption of zero); 3 streams.
func (c *Ctx) Encrypt(seed, stream uint64, level int, pt, pk0, pk1, c0, c1 *Poly) {
	var h C.mkhe_poly
	if pt != nil {
		h = pt.h
	}
	c.check(C.mkhe_encrypt(c.p, C.uint64_t(seed), C.uint64_t(stream), C.int(level), h, pk0.h, pk1.h, c0.h, c1.h))
}
