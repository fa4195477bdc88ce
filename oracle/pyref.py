"""pyref.py -- TEST INFRASTRUCTURE. Independent Python restatement of mdBRIEFextractorOct::operator()
that calls the *real* OpenCV (cv2) for every OpenCV primitive the reference calls, and restates only
the reference-specific logic.  It exists to pin oracle/mcs_oracle.cpp (which restates the OpenCV
arithmetic by hand) and to generate tests/golden/*.npz.  Needs cv2; only ever run in the build
container (oracle/pin_cv2.py), never on the GPU box.

Reference: /root/reference/src/mdBRIEFextractorOct.cpp (line numbers in the comments).
"""
import math
import numpy as np
import cv2

EDGE = 25
HALF_PATCH = 16
PATCH = 32
DEG2RADf = np.float32(np.float32(np.pi) / np.float32(180.0))
RHOd = 180.0 / 3.1415926535897932384626433832795028841971693993
RHOf = np.float32(np.float32(180.0) / np.float32(3.1415926535897932384626))


def cv_round(v):
    return int(np.rint(v))


def load_pairs():
    import pathlib
    p = pathlib.Path(__file__).resolve().parents[1] / "multicol_slam_b200/data/brief_pairs_64.bin"
    return np.frombuffer(p.read_bytes(), dtype=np.int8).astype(np.int64).reshape(-1, 2)


class Cam:
    def __init__(self, c, d, e, u0, v0, pol, inv_pol, width, height, mirror_mask=1):
        self.c, self.d, self.e, self.u0, self.v0 = c, d, e, u0, v0
        self.pol = list(pol) + [0.0] * (5 - len(pol))
        self.inv_pol = list(inv_pol) + [0.0] * (12 - len(inv_pol))
        self.width, self.height, self.mirror_mask = width, height, mirror_mask

    @staticmethod
    def horner(c, x):
        r = 0.0
        for v in reversed(c):
            r = r * x + v
        return r

    def world_to_img(self, x, y, z):      # src/cam_model_omni.cpp:146-161
        norm = math.sqrt(x * x + y * y)
        if norm == 0.0:
            norm = 1e-14
        theta = math.atan(-z / norm)
        rho = self.horner(self.inv_pol, theta)
        uu = x / norm * rho
        vv = y / norm * rho
        return uu * self.c + vv * self.d + self.u0, uu * self.e + vv + self.v0

    def img_to_world(self, u, v):          # src/cam_model_omni.cpp:49-67
        inv_aff = self.c - self.d * self.e
        u_t = u - self.u0
        v_t = v - self.v0
        x = (u_t - self.d * v_t) / inv_aff
        y = (-self.e * u_t + self.c * v_t) / inv_aff
        X2, Y2 = x * x, y * y
        z = -self.horner(self.pol, math.sqrt(X2 + Y2))
        n = math.sqrt(X2 + Y2 + z * z)
        return x / n, y / n, z / n

    def undistort(self, px, py, s):        # include/cam_model_omni.h:127-138
        x, y, z = self.img_to_world(px, py)
        return -x / z * s, -y / z * s

    def mirror_mask_img(self):             # src/cam_model_omni.cpp:181-220 (level 0)
        h, w = self.height, self.width
        if self.mirror_mask != 1:
            return np.ones((h, w), np.uint8)
        u0 = np.float32(self.v0)
        v0 = np.float32(self.u0)
        i = np.arange(h, dtype=np.float32)[:, None]
        j = np.arange(w, dtype=np.float32)[None, :]
        a = ((i - u0) ** 2).astype(np.float32) + ((j - v0) ** 2).astype(np.float32)
        ans = np.sqrt(a.astype(np.float32)).astype(np.float32)
        return np.where(ans < np.float32(u0 + np.float32(22.0)), 255, 0).astype(np.uint8)


class Extractor:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, fast_threshold=20, do_dbrief=False,
                 learn_masks=False, desc_size=32):
        self.nfeatures, self.nlevels, self.fast_threshold = nfeatures, nlevels, fast_threshold
        self.do_dbrief, self.learn_masks, self.desc_size = do_dbrief, learn_masks, desc_size
        sf = float(np.float32(scale_factor))
        self.scale_factor = sf
        self.sf = [1.0]
        for i in range(1, nlevels):
            self.sf.append(self.sf[-1] * sf)
        inv = 1.0 / sf
        self.isf = [1.0]
        for i in range(1, nlevels):
            self.isf.append(self.isf[-1] * inv)
        factor = 1.0 / sf
        nd = nfeatures * (1 - factor) / (1 - math.pow(factor, nlevels))
        self.quota, s = [], 0
        for l in range(nlevels - 1):
            self.quota.append(cv_round(nd))
            s += self.quota[-1]
            nd *= factor
        self.quota.append(max(nfeatures - s, 0))
        self.pattern = load_pairs()[:16 * desc_size]
        umax = [0] * (HALF_PATCH + 1)
        s2 = float(np.float32(HALF_PATCH) * np.sqrt(np.float32(2.0)) / np.float32(2))
        vmax = int(math.floor(s2 + 1))
        vmin = int(math.ceil(s2))
        for v in range(vmax + 1):
            umax[v] = cv_round(math.sqrt(HALF_PATCH * HALF_PATCH - v * v))
        v0 = 0
        for v in range(HALF_PATCH, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax

    # :1158-1201
    def compute_pyramid(self, image, mask):
        self.pyr, self.mpyr, self.roi = [], [], []
        for l in range(self.nlevels):
            scale = self.isf[l]
            w, h = cv_round(image.shape[1] * scale), cv_round(image.shape[0] * scale)
            if l != 0:
                pw, ph = self.roi[l - 1]
                prev = self.pyr[l - 1][EDGE:EDGE + ph, EDGE:EDGE + pw]
                prevm = self.mpyr[l - 1][EDGE:EDGE + ph, EDGE:EDGE + pw]
                cur = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
                curm = cv2.resize(prevm, (w, h), interpolation=cv2.INTER_NEAREST)
            else:
                cur, curm = image, mask
            self.pyr.append(cv2.copyMakeBorder(cur, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101))
            self.mpyr.append(cv2.copyMakeBorder(curm, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_CONSTANT, value=0))
            self.roi.append((w, h))

    # :863-949
    def detect_level(self, l):
        w, h = self.roi[l]
        img = self.pyr[l][EDGE:EDGE + h, EDGE:EDGE + w]
        msk = self.mpyr[l][EDGE:EDGE + h, EDGE:EDGE + w]
        fd = cv2.FastFeatureDetector_create(self.fast_threshold, True, 2)
        minB = EDGE - 3
        maxBX, maxBY = w - EDGE + 3, h - EDGE + 3
        width, height = float(maxBX - minB), float(maxBY - minB)
        nCols, nRows = int(width / 30.0), int(height / 30.0)
        wCell, hCell = int(math.ceil(width / nCols)), int(math.ceil(height / nRows))
        out = []
        for i in range(nRows):
            iniY = minB + i * hCell
            maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minB + j * wCell
                maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                kps = fd.detect(img[iniY:maxY, iniX:maxX], msk[iniY:maxY, iniX:maxX])
                for k in kps:
                    out.append((np.float32(k.pt[0]) + np.float32(j * wCell), np.float32(k.pt[1]) + np.float32(i * hCell),
                                np.float32(k.response)))
        return out

    # :569-861 ; python lists emulate std::list (index 0 == front)
    def distribute_octree(self, keys, minX, maxX, minY, maxY, N):
        class Node:
            __slots__ = ("keys", "UL", "UR", "BL", "BR", "no_more", "seq")
        seq = [0]

        def mk(UL, UR, BL, BR):
            n = Node()
            n.keys, n.UL, n.UR, n.BL, n.BR, n.no_more = [], UL, UR, BL, BR, False
            n.seq = seq[0]
            seq[0] += 1
            return n

        def divide(n):
            halfX = int(math.ceil((n.UR[0] - n.UL[0]) / 2.0))
            halfY = int(math.ceil((n.BR[1] - n.UL[1]) / 2.0))
            n1 = mk(n.UL, (n.UL[0] + halfX, n.UL[1]), (n.UL[0], n.UL[1] + halfY), (n.UL[0] + halfX, n.UL[1] + halfY))
            n2 = mk(n1.UR, n.UR, n1.BR, (n.UR[0], n.UL[1] + halfY))
            n3 = mk(n1.BL, n1.BR, n.BL, (n1.BR[0], n.BL[1]))
            n4 = mk(n3.UR, n2.BR, n3.BR, n.BR)
            for kp in n.keys:
                if kp[0] < n1.UR[0]:
                    (n1 if kp[1] < n1.BR[1] else n3).keys.append(kp)
                elif kp[1] < n1.BR[1]:
                    n2.keys.append(kp)
                else:
                    n4.keys.append(kp)
            for c in (n1, n2, n3, n4):
                if len(c.keys) == 1:
                    c.no_more = True
            return n1, n2, n3, n4

        nIni = cv_round(float(maxX - minX) / (maxY - minY))
        hX = float(maxX - minX) / nIni
        nodes = []
        for i in range(nIni):
            nodes.append(mk((int(hX * i), 0), (int(hX * (i + 1)), 0), (int(hX * i), maxY - minY),
                            (int(hX * (i + 1)), maxY - minY)))
        ini = list(nodes)
        for kp in keys:
            ini[int(float(kp[0]) / hX)].keys.append(kp)
        nn = []
        for n in nodes:
            if len(n.keys) == 1:
                n.no_more = True
                nn.append(n)
            elif len(n.keys) > 0:
                nn.append(n)
        nodes = nn
        finish = False
        while not finish:
            prev_size = len(nodes)
            n_to_expand = 0
            size_ptr = []
            front, keep = [], []
            for n in nodes:
                if n.no_more:
                    keep.append(n)
                    continue
                for c in divide(n):
                    if len(c.keys) > 0:
                        front.insert(0, c)
                        if len(c.keys) > 1:
                            n_to_expand += 1
                            size_ptr.append(c)
            nodes = front + keep
            if len(nodes) >= N or len(nodes) == prev_size:
                finish = True
            elif len(nodes) + n_to_expand * 3 > N:
                while not finish:
                    prev_size = len(nodes)
                    prev = sorted(size_ptr, key=lambda c: (len(c.keys), c.seq))
                    size_ptr = []
                    for n in reversed(prev):
                        for c in divide(n):
                            if len(c.keys) > 0:
                                nodes.insert(0, c)
                                if len(c.keys) > 1:
                                    size_ptr.append(c)
                        nodes.remove(n)
                        if len(nodes) >= N:
                            break
                    if len(nodes) >= N or len(nodes) == prev_size:
                        finish = True
        res = []
        for n in nodes:
            best = n.keys[0]
            for k in n.keys[1:]:
                if k[2] > best[2]:
                    best = k
            res.append(best)
        return res

    # :221-248
    def ic_angle(self, l, x, y):
        buf = self.pyr[l]
        cy, cx = cv_round(y) + EDGE, cv_round(x) + EDGE
        m01 = m10 = 0
        row = buf[cy].astype(np.int64)
        for u in range(-HALF_PATCH, HALF_PATCH + 1):
            m10 += u * int(row[cx + u])
        for v in range(1, HALF_PATCH + 1):
            d = self.umax[v]
            rp = buf[cy + v, cx - d:cx + d + 1].astype(np.int64)
            rm = buf[cy - v, cx - d:cx + d + 1].astype(np.int64)
            u = np.arange(-d, d + 1)
            m01 += v * int((rp - rm).sum())
            m10 += int((u * (rp + rm)).sum())
        return np.float32(cv2.fastAtan2(float(np.float32(m01)), float(np.float32(m10))))

    def rotate_pattern(self, ax, ay):            # :285-301
        return [(cv_round(int(px) * ax - int(py) * ay), cv_round(int(px) * ay + int(py) * ax)) for px, py in self.pattern]

    def rotate_distort(self, ukx, uky, cam, ax, ay):   # :250-283
        xs, ys = [], []
        sx = sy = 0.0
        for px, py in self.pattern:
            xr = int(px) * ax - int(py) * ay + ukx
            yr = int(px) * ay + int(py) * ax + uky
            u, v = cam.world_to_img(xr, yr, -cam.pol[0])
            xs.append(u)
            ys.append(v)
            sx += u
            sy += v
        n = float(len(xs))
        mx, my = sx / n, sy / n
        return [(cv_round(x - mx), cv_round(y - my)) for x, y in zip(xs, ys)]

    def describe(self, blurred, kp, uk, cam):    # :303-554
        x, y, angle = kp
        row, col = cv_round(y), cv_round(x)
        if self.learn_masks:
            rot = 20.0 / RHOd
            a = float(np.float32(angle) / RHOf)
            pat = self.rotate_distort(uk[0], uk[1], cam, math.cos(a), math.sin(a))
            m1 = self.rotate_distort(uk[0], uk[1], cam, math.cos(a + rot), math.sin(a + rot))
            m2 = self.rotate_distort(uk[0], uk[1], cam, math.cos(a - rot), math.sin(a - rot))
        else:
            a = float(np.float32(angle) * DEG2RADf)
            if self.do_dbrief:
                pat = self.rotate_distort(uk[0], uk[1], cam, math.cos(a), math.sin(a))
            else:
                pat = self.rotate_pattern(math.cos(a), math.sin(a))

        def S(p):
            return int(blurred[row + p[1] + EDGE, col + p[0] + EDGE])
        desc = np.zeros(self.desc_size, np.uint8)
        dm = np.zeros(self.desc_size, np.uint8)
        for i in range(self.desc_size):
            val = mval = 0
            for b in range(8):
                k = 16 * i + 2 * b
                t = int(S(pat[k]) < S(pat[k + 1]))
                val |= t << b
                if self.learn_masks:
                    s1 = int(S(m1[k]) < S(m1[k + 1])) ^ t
                    s2 = int(S(m2[k]) < S(m2[k + 1])) ^ t
                    mval |= int(s1 + s2 == 0) << b
            desc[i], dm[i] = val, mval
        return desc, dm

    # :1244-1337 ; returns dict of everything (intermediates included)
    def __call__(self, image, mask, cam):
        self.compute_pyramid(image, mask)
        allk, raws = [], []
        for l in range(self.nlevels):
            raw = self.detect_level(l)
            raws.append(raw)
            w, h = self.roi[l]
            minB = EDGE - 3
            sel = self.distribute_octree(raw, minB, w - EDGE + 3, minB, h - EDGE + 3, self.quota[l]) if raw else []
            allk.append([(np.float32(k[0] + np.float32(minB)), np.float32(k[1] + np.float32(minB)), k[2]) for k in sel])
        angles = [[self.ic_angle(l, k[0], k[1]) for k in allk[l]] for l in range(self.nlevels)]
        kps, descs, masks = [], [], []
        blurred_levels = []
        for l in range(self.nlevels):
            w, h = self.roi[l]
            buf = self.pyr[l].copy()
            if allk[l]:
                roi = buf[EDGE:EDGE + h, EDGE:EDGE + w]
                cv2.boxFilter(roi, -1, (5, 5), roi, (-1, -1), True, cv2.BORDER_REFLECT_101)
            blurred_levels.append(buf)
            scale = np.float32(self.sf[l])
            for k, a in zip(allk[l], angles[l]):
                uk = (0.0, 0.0)
                if self.do_dbrief:
                    uk = cam.undistort(float(np.float32(k[0] * scale)), float(np.float32(k[1] * scale)), cam.pol[0])
                d, m = self.describe(buf, (k[0], k[1], a), uk, cam)
                ox, oy = (k[0], k[1]) if l == 0 else (np.float32(k[0] * scale), np.float32(k[1] * scale))
                kps.append((ox, oy, np.float32(int(PATCH * self.sf[l])), a, k[2], l, -1))
                descs.append(d)
                masks.append(m)
        return dict(kps=kps, desc=np.array(descs, np.uint8).reshape(-1, self.desc_size),
                    dmask=np.array(masks, np.uint8).reshape(-1, self.desc_size), raws=raws,
                    pyr=[self.pyr[l][EDGE:EDGE + self.roi[l][1], EDGE:EDGE + self.roi[l][0]] for l in range(self.nlevels)],
                    blur=[blurred_levels[l][EDGE:EDGE + self.roi[l][1], EDGE:EDGE + self.roi[l][0]] for l in range(self.nlevels)],
                    mpyr=[self.mpyr[l][EDGE:EDGE + self.roi[l][1], EDGE:EDGE + self.roi[l][0]] for l in range(self.nlevels)])
