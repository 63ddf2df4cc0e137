"""Upper bound on what spatial splits could buy on the alpha-tested foliage scene (config 4), measured on the CPU through the reference traverser's
own counters: every card triangle is cut along a regular grid of cell size h and the PIECES are handed to the product builder as triangles
(what an SBVH with unlimited reference duplication would enclose).  Same card density as the full-size config (2.1 cards per unit volume).
usage: python tools/sbvh_probe.py"""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import vistrace_b200 as vt, oracle
from vistrace_b200 import scenes, abi
f4=np.float32
E=20.0
scene = scenes.scene_foliage(n_cards=20000, extent=E, tex_size=256, ground_quads=8)
rays = scenes.pinhole_rays(320, 180, (0, -E*0.48, 20), (0, 0, 8))
kind = "reference" if oracle.available("reference") else "port"
def stats(sc, label):
    nodes, prims = vt.build_bvh(sc)
    cpu = oracle.CpuScene(sc, kind, build_bvh=False); cpu.set_bvh(nodes, prims)
    out = cpu.traverse(rays, want_stats=True)
    n=len(rays)
    print(f"{label}: tris {len(sc.tris)} nodes {len(nodes)} steps/ray {out['steps']/n:.1f} tests/ray {out['isects']/n:.1f} hits {(out['hits']['prim']!=0xFFFFFFFF).mean():.3f}", flush=True)
    return out
base = stats(scene, "unsplit")

def clip_poly(P, axis, val, keep_less):
    # P: list of (pos(3), bary(3)); Sutherland-Hodgman against plane x[axis]=val
    out=[]
    n=len(P)
    for i in range(n):
        a=P[i]; b=P[(i+1)%n]
        da=a[0][axis]-val; db=b[0][axis]-val
        ina = da<=0 if keep_less else da>=0
        inb = db<=0 if keep_less else db>=0
        if ina: out.append(a)
        if ina!=inb:
            t=da/(da-db)
            out.append((a[0]+(b[0]-a[0])*t, a[1]+(b[1]-a[1])*t))
    return out

def split_scene(sc, h):
    T=sc.tris
    nfol = int((T["material"]>=2).sum())
    newt=[]
    t0=time.time()
    for ti in range(len(T)):
        t=T[ti]
        if t["material"]<2:
            newt.append(t.copy()); continue
        P=[(t["p"][k].astype(np.float64), np.eye(3)[k]) for k in range(3)]
        polys=[P]
        for axis in range(3):
            nxt=[]
            for poly in polys:
                lo=min(v[0][axis] for v in poly); hi=max(v[0][axis] for v in poly)
                k0=int(np.floor(lo/h)); k1=int(np.floor(hi/h))
                cur=poly
                for k in range(k0,k1+1):
                    piece=clip_poly(cur, axis, (k+1)*h, True) if k<k1 else cur
                    if len(piece)>=3: nxt.append(piece)
                    if k<k1:
                        cur=clip_poly(cur, axis, (k+1)*h, False)
                        if len(cur)<3: break
            polys=nxt
        for poly in polys:
            for j in range(1,len(poly)-1):
                tri=t.copy()
                idx=(0,j,j+1)
                for k,q in enumerate(idx):
                    b=poly[q][1]
                    tri["p"][k]=poly[q][0].astype(f4)
                    tri["uvs"][k]=(b[:,None]*t["uvs"]).sum(0).astype(f4)
                    tri["normals"][k]=(b[:,None]*t["normals"]).sum(0).astype(f4)
                    tri["alphas"][k]=float((b*t["alphas"]).sum())
                # skip degenerate
                a=np.cross(tri["p"][1]-tri["p"][0], tri["p"][2]-tri["p"][0])
                if np.dot(a,a)>1e-12: newt.append(tri)
    out=np.array(newt, dtype=abi.TRI_IN)
    print(f"split h={h}: {len(T)} -> {len(out)} tris in {time.time()-t0:.0f}s", flush=True)
    return abi.SceneData(out, sc.materials, sc.entities, [(w,hh,m,fl,px) for (w,hh,m,fl,px,lay) in sc.textures])
for h in (1.5, 0.8):
    s2=split_scene(scene,h)
    o=stats(s2, f"pieces h={h}")
    d=np.abs(o['hits']['t']-base['hits']['t']); print("  max |dt|", float(d.max()), "hit/miss diff", int(((o['hits']['prim']!=0xFFFFFFFF)!=(base['hits']['prim']!=0xFFFFFFFF)).sum()))
