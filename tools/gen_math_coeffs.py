"""Generates the coefficients used by csrc/gabo_math.cuh (run once; output pasted into the header).
  asin(x) = x + x*z*P(z), z = x^2 in [0, 0.25]  (near-minimax: Chebyshev interpolation in 60-digit arithmetic)
  exp(r)  = sum r^k/k!,  |r| <= ln2/64, degree 6, with a 32-entry table of 2^(j/32)."""
import mpmath as mp
mp.mp.dps = 60

def g(z):
    if z == 0:
        return mp.mpf(1) / 6
    s = mp.sqrt(z)
    return (mp.asin(s) / s - 1) / z

def cheb_fit(f, a, b, n):
    # interpolate at Chebyshev nodes, return monomial coefficients in z
    nodes = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * (n + 1))) for k in range(n + 1)]
    V = mp.matrix(n + 1, n + 1)
    y = mp.matrix(n + 1, 1)
    for i, x in enumerate(nodes):
        for j in range(n + 1):
            V[i, j] = x ** j
        y[i] = f(x)
    return list(mp.lu_solve(V, y))

for deg in (10, 11, 12, 13):
    c = cheb_fit(g, mp.mpf(0), mp.mpf('0.25'), deg)
    cd = [float(v) for v in c]
    # error of acos built from double coefficients (evaluated in mp with double coeffs): both branches
    worst = 0
    for k in range(2001):
        t = mp.mpf(-1) + mp.mpf(2) * k / 2000
        a = abs(t)
        if a <= 0.5:
            z = t * t; p = z * mp.polyval(cd[::-1], z); d = mp.pi / 2 - (t + t * p)
        else:
            z = (1 - a) / 2; s = mp.sqrt(z); p = z * mp.polyval(cd[::-1], z); r = s + s * p
            d = 2 * r if t > 0 else mp.pi - 2 * r
        worst = max(worst, abs(d - mp.acos(t)))
    print('deg', deg, 'max abs err of acos', mp.nstr(worst, 3))
    if deg == 11:
        best = cd
print('static constexpr double kAsinP[] = {')
for v in best:
    print('    %s,' % float(v).hex())
print('};')
print('// exp table 2^(j/32)')
print('static const double kExp2Tab[32] = {')
for j in range(32):
    print('    %s,' % float(mp.mpf(2) ** (mp.mpf(j) / 32)).hex())
print('};')
ln2_32 = mp.log(2) / 32
hi = float(ln2_32)
# make hi have trailing zero bits so n*hi is exact for |n| < 2^16: keep 36 bits
import struct
bits = struct.unpack('<Q', struct.pack('<d', hi))[0] & ~((1 << 17) - 1)
hi = struct.unpack('<d', struct.pack('<Q', bits))[0]
lo = float(ln2_32 - mp.mpf(hi))
print('L2E32 =', float(32 / mp.log(2)).hex(), ' LN2_32_HI =', hi.hex(), ' LN2_32_LO =', lo.hex())
print('pi/2 =', float(mp.pi / 2).hex(), ' pi =', float(mp.pi).hex())
