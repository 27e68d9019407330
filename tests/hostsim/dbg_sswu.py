import ctypes, sys, os
sys.path.insert(0, '/root/repo' if os.path.exists('/root/repo/oracle') else '.')
sys.path.insert(0, os.getcwd())
from oracle import bls_oracle as O
libname = sys.argv[1]
hs = ctypes.CDLL(libname)
p = O.p
def b48(v): return (v % p).to_bytes(48, 'big')
def fp2b(a): return b48(a[0]) + b48(a[1])
def bfp2(b): return (int.from_bytes(b[:48], 'big'), int.from_bytes(b[48:96], 'big'))
u = (1720201260439466020000 + 12345, 78018059744264389400 + 999)
out = ctypes.create_string_buffer(1000)
hs.hs_sswu_dbg(fp2b(u), out)
r = out.raw
tv1 = O.f2_mul(O.f2_sqr(u), O.SSWU_Z)
tv2 = O.f2_add(O.f2_sqr(tv1), tv1)
xn = O.f2_neg(O.f2_mul(O.f2_add(tv2, (1,0)), O.SSWU_B)); xd = O.f2_mul(tv2, O.SSWU_A)
gxd = O.f2_mul(O.f2_sqr(xd), xd)
gxn = O.f2_add(O.f2_mul(O.f2_add(O.f2_sqr(xn), O.f2_mul(O.f2_sqr(xd), O.SSWU_A)), xn), O.f2_mul(gxd, O.SSWU_B))
n = (gxd[0]**2 + gxd[1]**2) % p
wv = O.f2_muls(O.f2_mul(gxn, O.f2_conj(gxd)), n)
print('tv1', bfp2(r[0:96]) == tv1)
print('gxn', bfp2(r[96:192]) == gxn)
print('gxd', bfp2(r[192:288]) == gxd)
print('n', int.from_bytes(r[288:336],'big') == n)
print('wv', bfp2(r[336:432]) == wv)
root = bfp2(r[432:528]); sq = r[672]
print('sq', sq, 'root^2==wv', O.f2_sqr(root) == wv, 'root^2==Z wv', O.f2_sqr(root) == O.f2_mul(O.SSWU_Z, wv))
ninv = int.from_bytes(r[528:576],'big'); print('ninv', ninv == pow(n, -1, p))
root2 = bfp2(r[576:672]); print('root2', root2 == O.f2_muls(root, ninv))
print('sgn', r[673], O.f2_sgn0(u), r[674], O.f2_sgn0(root2))

root3 = bfp2(r[700:796]); exp3 = O.f2_mul(root2, O.f2_mul(tv1, u)) if not sq else root2
print('root3', root3 == exp3, 'sgn3', r[675], O.f2_sgn0(root3))
root4 = bfp2(r[800:896]); exp4 = root3 if O.f2_sgn0(u) == O.f2_sgn0(root3) else O.f2_neg(root3)
print('root4', root4 == exp4)
ry = bfp2(r[900:996]); print('sswu_g2 y == inline', ry == root4, ' == oracle', ry == O.simplified_swu_fp2(u)[1])

print('variants A,B,neg, ry==root4, rxn, rxd:', list(r[676:682]))
cgxd = O.f2_conj(gxd)
t_kar = O.f2_mul(tv1, u)
cands = {'root': root, 'root2': root2, 'root3': root3, '-root3': O.f2_neg(root3), 'root2*conj(gxd)': O.f2_mul(root2, cgxd),
         'root*t': O.f2_mul(root, t_kar), 'root2*tv1': O.f2_mul(root2, tv1), 'root2*u': O.f2_mul(root2, u),
         'root2*xn': O.f2_mul(root2, xn), 'root2*xn2': O.f2_mul(root2, O.f2_mul(xn, tv1))}
for k, v in cands.items():
    for sgn, vv in (('+', v), ('-', O.f2_neg(v))):
        if vv == ry: print('MATCH', sgn, k)
        if vv[0] == ry[0]: print('re match', sgn, k)
        if vv[1] == ry[1]: print('im match', sgn, k)
print('ry', hex(ry[0])[:20], hex(ry[1])[:20]); print('root4', hex(root4[0])[:20], hex(root4[1])[:20])
print('ry^2 == gx2?', O.f2_sqr(ry) == O.f2_mul(O.f2_sqr(root4), (1,0)))

print('inside sswu_g2: root,root2,root3,root4 ok; dbg4==ry:', list(r[682:687]))
