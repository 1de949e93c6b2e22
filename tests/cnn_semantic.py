"""Semantic end-to-end check of the encrypted CNN (BASELINE config 5) on material made entirely on the device.

The packing of image, kernels, FC1 / FC2 matrices and biases into slots restates the encrypt* helpers of the reference's test
(cnn/cnn_test.go:345-545: test infrastructure there as well); weights and image are random (the reference's mnist_test.csv is
not shipped, and what is checked here is the arithmetic, not the accuracy of a trained model): the logits that come out of
    device key generation -> device encryption -> cnn.inference on the device -> device decryption -> decode
must equal the plain forward pass  conv(4x4, stride 2, 5 kernels) -> x^2 -> FC1 + B1 -> x^2 -> FC2 + B2  within CKKS precision.
This pins what the op sequence MEANS (rotation directions, slot layout, scale management), which bit-equality with the oracle on
random ciphertexts (tests/cnn_flow.py) cannot."""
import numpy as np

IMAGE, NCLASS, NKERN, KSIZE, BLOCK, CONV, NFC, GAP = 28, 10, 5, 4, 14, 13, 64, 128      # cnn/cnn_test.go:22-33


def pack_image(img):
    """encryptImage, cnn_test.go:345-377: the four stride-2 phases of the image, one copy per kernel, then mirrored into the upper half"""
    v = np.zeros(8192)
    for k in range(NKERN):
        for i in range(BLOCK):
            for j in range(BLOCK):
                idx = BLOCK * BLOCK * k + BLOCK * i + j
                v[idx] = img[2 * i][2 * j]
                v[idx + 1024] = img[2 * i][2 * j + 1]
                v[idx + 2048] = img[2 * i + 1][2 * j]
                v[idx + 3072] = img[2 * i + 1][2 * j + 1]
    v[4096:] = v[:4096]
    return v


def pack_kernels(kern):
    """encryptKernels, cnn_test.go:379-441: ciphertext m = (p', q') holds kernel entry (2p' + a, 2q' + b) in phase block (a, b)"""
    out = np.zeros((4, 8192))
    for i in range(NKERN):
        for j in range(CONV):
            for k in range(CONV):
                idx = BLOCK * BLOCK * i + BLOCK * j + k
                for m, (p, q) in enumerate(((0, 0), (0, 2), (2, 0), (2, 2))):
                    out[m][idx] = kern[i][p][q]
                    out[m][idx + 1024] = kern[i][p][q + 1]
                    out[m][idx + 2048] = kern[i][p + 1][q]
                    out[m][idx + 3072] = kern[i][p + 1][q + 1]
    out[:, 4096:] = out[:, :4096]
    return out


def pack_fc1(FC1):
    """encryptFC1, cnn_test.go:443-486: FC1 is (5*13*13) x 64; diagonal packing over 8 ciphertexts of 64 x 128 slots"""
    temp = np.zeros((NFC, 1024))
    for i in range(NKERN):
        for j in range(CONV):
            for k in range(CONV):
                temp[:, BLOCK * BLOCK * i + BLOCK * j + k] = FC1[i + NKERN * (j * CONV + k)]
    out = np.zeros((8, 8192))
    for i in range(8):
        for j in range(64):
            out[i][128 * j:128 * j + 128] = temp[j][128 * ((i + j) % 8):128 * ((i + j) % 8) + 128]
    return out


def pack_fc2(FC2):
    """encryptFC2, cnn_test.go:488-511: slot 128 x + y = FC2[x][y]"""
    v = np.zeros(8192)
    for x in range(NFC):
        v[x * GAP:x * GAP + NCLASS] = FC2[x]
    return v


def pack_b1(B1):
    v = np.zeros(8192)
    v[np.arange(NFC) * GAP] = B1
    return v


def pack_b2(B2):
    v = np.zeros(8192)
    v[:NCLASS] = B2
    return v


def plain_forward(img, kern, FC1, FC2, B1, B2):
    conv = np.zeros((NKERN, CONV, CONV))
    for k in range(NKERN):
        for j in range(CONV):
            for l in range(CONV):
                conv[k, j, l] = np.sum(img[2 * j:2 * j + KSIZE, 2 * l:2 * l + KSIZE] * kern[k])
    h = conv ** 2
    x = np.array([h[i, j, k] for j in range(CONV) for k in range(CONV) for i in range(NKERN)])      # FC1 row index i + 5 (13 j + k)
    f1 = (x @ FC1 + B1) ** 2
    return f1 @ FC2 + B2


def run(lit, lib=None, seed=3):
    """returns (logits decrypted from the device run, logits of the plain forward pass)"""
    from oracle import oracle as O
    from mkhe_kklss_b200 import cnn, mkckks, mkrlwe
    assert lit.logN == 14, "the reference's packing fills 8192 slots"
    rng = np.random.default_rng(seed)
    img = rng.uniform(0, 1, (IMAGE, IMAGE))
    kern = rng.uniform(-0.5, 0.5, (NKERN, KSIZE, KSIZE))
    FC1 = rng.uniform(-0.1, 0.1, (NKERN * CONV * CONV, NFC))
    FC2 = rng.uniform(-0.2, 0.2, (NFC, NCLASS))
    B1, B2 = rng.uniform(-0.2, 0.2, NFC), rng.uniform(-0.2, 0.2, NCLASS)
    want = plain_forward(img, kern, FC1, FC2, B1, B2)

    op = O.MKParams(lit.logN, lit.Q, lit.P, lit.gamma, crs_rots=[])           # host side: encoder / decoder only
    dp = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale, lib=lib, gamma=lit.gamma)
    try:
        rots = [r for r in cnn.cnn_rotations(lit.logN) if r < lit.N // 2]
        SEED = 0xC0FFEE
        for idx in [0, -1] + rots:
            dp.AddCRS(idx, SEED)
        kg = mkrlwe.KeyGenerator(dp, SEED + 1)
        rl, rk, sk, pk = mkrlwe.RelinearizationKeySet(), mkrlwe.RotationKeySet(), {}, {}
        for i in (cnn.MODEL, cnn.DATA):
            sk[i], pk[i] = kg.GenKeyPair(i)
            rl.AddRelinearizationKey(kg.GenRelinearizationKey(sk[i], kg.GenSecretKey(i)))
            for r in rots:
                rk.AddRotationKey(i, r, kg.GenRotationKey(r, sk[i]))
        enc = mkrlwe.Encryptor(dp, SEED + 2)
        L = op.max_level()

        def encrypt(slots, owner):
            pt = mkrlwe.Poly.from_numpy(dp.ctx, O.ckks_encode(op, slots.astype(np.complex128), lit.scale, L))
            ct = mkckks.Ciphertext.new(dp, [owner], L, lit.scale)
            enc.Encrypt(pt, pk[owner], ct)
            return ct

        ctImage = encrypt(pack_image(img), cnn.DATA)
        ctKernels = [encrypt(v, cnn.MODEL) for v in pack_kernels(kern)]
        ctFC1 = [encrypt(v, cnn.MODEL) for v in pack_fc1(FC1)]
        ctFC2, ctB1, ctB2 = encrypt(pack_fc2(FC2), cnn.MODEL), encrypt(pack_b1(B1), cnn.MODEL), encrypt(pack_b2(B2), cnn.MODEL)
        mask = np.zeros(8192)
        mask[::GAP] = 1                                                        # cnn_test.go:141-148
        ptMask = mkrlwe.Poly.from_numpy(dp.ctx, O.ckks_encode(op, mask.astype(np.complex128), lit.scale, L))
        E = cnn.DeviceFacade(mkckks.Evaluator(dp), rl, rk)
        out, _ = cnn.inference(E, ctImage, ctKernels, ctFC1, ctFC2, ctB1, ctB2, ptMask, lit.scale)
        pt = mkrlwe.Decryptor(dp).Decrypt(out, sk).numpy()
        got = O.ckks_decode(op, pt, out.Scale)[:NCLASS].real
        return got, want
    finally:
        dp.ctx.close()


def run_oracle(lit, seed=3):
    """the same experiment on the CPU oracle alone (keys, encryption, inference, decryption by oracle/): pins the oracle's op
    sequence semantically, like the reference's TestCNN pins the Go evaluator (cnn/cnn_test.go:97-166)"""
    from oracle import oracle as O
    from mkhe_kklss_b200 import cnn
    import cnn_flow as F
    assert lit.logN == 14
    rng = np.random.default_rng(seed)
    img = rng.uniform(0, 1, (IMAGE, IMAGE))
    kern = rng.uniform(-0.5, 0.5, (NKERN, KSIZE, KSIZE))
    FC1 = rng.uniform(-0.1, 0.1, (NKERN * CONV * CONV, NFC))
    FC2 = rng.uniform(-0.2, 0.2, (NFC, NCLASS))
    B1, B2 = rng.uniform(-0.2, 0.2, NFC), rng.uniform(-0.2, 0.2, NCLASS)
    want = plain_forward(img, kern, FC1, FC2, B1, B2)
    rots = [r for r in cnn.cnn_rotations(lit.logN) if r < lit.N // 2]
    op = O.MKParams(lit.logN, lit.Q, lit.P, lit.gamma, crs_rots=rots)
    kg = O.KeyGenerator(op)
    sk, pk, rl, rk = {}, {}, {}, {}
    for i in (cnn.MODEL, cnn.DATA):
        sk[i] = kg.gen_secret_key(i)
        pk[i] = kg.gen_public_key(sk[i])
        rl[i] = kg.gen_relin_key(sk[i], kg.gen_secret_key(i))
        rk[i] = {r: kg.gen_rotation_key(r, sk[i]) for r in rots}
    enc = O.Encryptor(op)
    L = op.max_level()
    encrypt = lambda slots, owner: enc.encrypt(O.ckks_encode(op, slots.astype(np.complex128), lit.scale, L), pk[owner], owner, lit.scale)
    ctImage = encrypt(pack_image(img), cnn.DATA)
    ctKernels = [encrypt(v, cnn.MODEL) for v in pack_kernels(kern)]
    ctFC1 = [encrypt(v, cnn.MODEL) for v in pack_fc1(FC1)]
    ctFC2, ctB1, ctB2 = encrypt(pack_fc2(FC2), cnn.MODEL), encrypt(pack_b1(B1), cnn.MODEL), encrypt(pack_b2(B2), cnn.MODEL)
    mask = np.zeros(8192)
    mask[::GAP] = 1
    ptMask = O.ckks_encode(op, mask.astype(np.complex128), lit.scale, L)
    E = F.OracleFacade(O.CKKSEvaluator(op, lit.scale), rl, rk)
    out, _ = cnn.inference(E, ctImage, ctKernels, ctFC1, ctFC2, ctB1, ctB2, ptMask, lit.scale)
    got = O.ckks_decode(op, O.Decryptor(op).decrypt(out, sk), out.scale)[:NCLASS].real
    return got, want
