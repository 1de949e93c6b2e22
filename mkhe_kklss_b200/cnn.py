"""Host-side mirror of the reference's encrypted-CNN evaluation (cnn/cnn.go:10-96, driven by cnn/cnn_test.go:121-162 and timed by
cnn/cnn_bench_test.go:12-73): Convolution, FC1Layer and FC2Layer written against a Go-named evaluator facade (BASELINE config 5).
The functions only issue evaluator calls -- every polynomial operation is a C-ABI call on the device; ids, levels, scales,
rotation indices and the order of operations are the reference's."""
import math

MODEL, DATA = 0, 1                       # modelOwner / dataOwner (cnn/cnn_test.go:35-36)
# genTestParams (cnn/cnn_test.go:185-201): extra CRS / rotation keys, plus every power of two below N/2
EXTRA_ROTS = [14, 15, 384, 512, 640, 768, 896, 8191, 8190, 8188, 8184]


def cnn_rotations(logN):
    return sorted(set(EXTRA_ROTS + [1 << i for i in range(logN - 1)]))

def convolution(E, ctImage, ctImageHoisted, ctKernels, ctKernelsHoisted):
    """Convolution, cnn/cnn.go:10-40"""
    convOut = E.MulRelinHoistedNew(ctImage, ctKernels[0], ctImageHoisted, ctKernelsHoisted[0])
    for j, rot in ((1, 1), (2, 14), (3, 15)):
        temp = E.RotateHoistedNew(ctImage, rot, ctImageHoisted)
        tempHoisted = E.HoistedForm(temp)
        temp = E.MulRelinHoistedNew(temp, ctKernels[j], tempHoisted, ctKernelsHoisted[j])
        convOut = E.AddNew(convOut, temp)
    for rot in (2048, 1024):
        temp = E.RotateNew(convOut, rot)
        convOut = E.AddNew(convOut, temp)
    return convOut


def fc1_layer(E, ctVec, ctVecHoisted, ctMat, ctMatHoisted, ctBias):
    """FC1Layer, cnn/cnn.go:42-71"""
    fc1Out = None
    for i in range(len(ctMat)):
        temp = E.RotateHoistedNew(ctVec, i * 128, ctVecHoisted)
        tempHoisted = E.HoistedForm(temp)
        temp = E.MulRelinHoistedNew(temp, ctMat[i], tempHoisted, ctMatHoisted[i])
        fc1Out = E.CopyNew(temp) if i == 0 else E.AddNew(fc1Out, temp)
    for i in range(int(math.log2(128))):
        temp = E.RotateNew(fc1Out, 1 << i)
        fc1Out = E.AddNew(fc1Out, temp)
    return E.AddNew(fc1Out, ctBias)


def fc2_layer(E, ctVec, ctMat, ctBias, ptMask, ptScale):
    """FC2Layer, cnn/cnn.go:73-96"""
    fc2Out = E.MulPtxtNew(ctVec, ptMask, ptScale)
    for i in range(int(math.log2(16))):
        temp = E.RotateNew(fc2Out, -1 * (1 << i))
        fc2Out = E.AddNew(fc2Out, temp)
    fc2Out = E.MulRelinNew(fc2Out, ctMat)
    for i in range(int(math.log2(64))):
        temp = E.RotateNew(fc2Out, 128 * (1 << i))
        fc2Out = E.AddNew(fc2Out, temp)
    return E.AddNew(fc2Out, ctBias)


def hoist_model(E, ctKernels, ctFC1):
    """the model owner's precomputation (cnn/cnn_bench_test.go:43-52): hoisted forms of the kernels and of the FC1 matrix"""
    return [E.HoistedForm(c) for c in ctKernels], [E.HoistedForm(c) for c in ctFC1]


def infer(E, ctImage, ctKernels, ctKernelsHoisted, ctFC1, ctFC1Hoisted, ctFC2, ctB1, ctB2, ptMask, ptScale):
    """one image through the network (cnn/cnn_bench_test.go:40,63-73 / cnn_test.go:130-162)"""
    ctImageHoisted = E.HoistedForm(ctImage)
    convOut = convolution(E, ctImage, ctImageHoisted, ctKernels, ctKernelsHoisted)
    convOutHoisted = E.HoistedForm(convOut)
    square1Out = E.MulRelinHoistedNew(convOut, convOut, convOutHoisted, convOutHoisted)
    square1OutHoisted = E.HoistedForm(square1Out)
    fc1Out = fc1_layer(E, square1Out, square1OutHoisted, ctFC1, ctFC1Hoisted, ctB1)
    fc1OutHoisted = E.HoistedForm(fc1Out)
    square2Out = E.MulRelinHoistedNew(fc1Out, fc1Out, fc1OutHoisted, fc1OutHoisted)
    return fc2_layer(E, square2Out, ctFC2, ctB2, ptMask, ptScale), {"conv": convOut, "square1": square1Out, "fc1": fc1Out}


def inference(E, ctImage, ctKernels, ctFC1, ctFC2, ctB1, ctB2, ptMask, ptScale):
    """TestCNN's evaluation, cnn/cnn_test.go:121-162"""
    kh, fh = hoist_model(E, ctKernels, ctFC1)
    return infer(E, ctImage, ctKernels, kh, ctFC1, fh, ctFC2, ctB1, ctB2, ptMask, ptScale)


class DeviceFacade:
    """mkckks.Evaluator of this package with the key sets bound, Go method names"""

    def __init__(self, ev, rlk, rk):
        self.ev, self.rlk, self.rk = ev, rlk, rk
        self.counts = {}

    def _c(self, n):
        self.counts[n] = self.counts.get(n, 0) + 1

    def HoistedForm(self, ct):
        self._c("HoistedForm"); return self.ev.HoistedForm(ct)

    def MulRelinHoistedNew(self, a, b, ha, hb):
        self._c("MulRelinHoistedNew"); return self.ev.MulRelinHoistedNew(a, b, ha, hb, self.rlk)

    def MulRelinNew(self, a, b):
        self._c("MulRelinNew"); return self.ev.MulRelinNew(a, b, self.rlk)

    def RotateHoistedNew(self, ct, rot, h):
        self._c("RotateHoistedNew"); return self.ev.RotateHoistedNew(ct, rot, h, self.rk)

    def RotateNew(self, ct, rot):
        self._c("RotateNew"); return self.ev.RotateNew(ct, rot, self.rk)

    def AddNew(self, a, b):
        self._c("AddNew"); return self.ev.AddNew(a, b)

    def CopyNew(self, ct):
        return self.ev.CopyNew(ct)

    def MulPtxtNew(self, ct, pt, scale):
        self._c("MulPtxtNew"); return self.ev.MulPtxtNew(ct, pt, scale)
