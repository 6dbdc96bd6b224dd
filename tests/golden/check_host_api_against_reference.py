"""Differential check of the host-side API (no kernels involved) against the UNMODIFIED reference, run in the build
container where /root/reference exists (python tests/golden/check_host_api_against_reference.py): every line must start
with SAME - same return values, same exception types and messages, same stdout.  Not a pytest test: the GPU boxes and
the fixtures-only tests never read /root/reference.  The reference's Cython modules and absent third-party imports
are stubbed; nothing here evaluates a residual."""
import sys, types, io, contextlib
import numpy as np
for name in ("numdifftools","matplotlib","matplotlib.pyplot","splitnewton","splitnewton.newton","splitnewton.split_newton"):
    sys.modules[name]=types.ModuleType(name)
sys.modules["splitnewton.newton"].newton=lambda *a,**k: None
sys.modules["splitnewton.split_newton"].split_newton=lambda *a,**k: None
for name, attrs in (("splitfxm.derivatives", ("Dx","D2x","dx","d2x")), ("splitfxm.flux", ("fluxes","diffusion_fluxes"))):
    m=types.ModuleType(name)
    for a in attrs: setattr(m,a,lambda *x,**k: None)
    sys.modules[name]=m
sys.path.insert(0,'/root/reference'); sys.path.insert(0,'/root/repo')
import splitfxm as R
from splitfxm.domain import Domain as RD
from splitfxm.simulation import Simulation as RS
from splitfxm.models.advection_diffusion import AdvectionDiffusion as RAD
from splitfxm.schemes import default_scheme as rdefault
from splitfxm.error import SFXM as RSFXM
import splitfxm_b200 as P
from splitfxm_b200.models.advection_diffusion import AdvectionDiffusion as PAD

comps=["u","v","w"]; ics={"u":"gaussian","v":"rarefaction","w":"tophat"}
bcs={"u":{"left":"periodic","right":"periodic"},"v":{"left":{"dirichlet":3},"right":{"dirichlet":4}},"w":{"left":{"dirichlet":2},"right":"periodic"}}
def both(fr, fp, label):
    out=[]
    for f in (fr, fp):
        try:
            buf=io.StringIO()
            with contextlib.redirect_stdout(buf):
                v=f()
            out.append(("ok", v, buf.getvalue()))
        except Exception as e:
            out.append(("err", type(e).__name__, str(e)))
    a,b=out
    def canon(v):
        if isinstance(v,(tuple,list)): return [canon(t) for t in v]
        if isinstance(v,np.ndarray): return v.tolist()
        if isinstance(v,(np.floating,np.integer,np.bool_)): return v.item()
        return v
    if a[0]=="ok" and b[0]=="ok": same = canon(a[1])==canon(b[1]) and a[2]==b[2]
    else: same = a==b
    print(("SAME " if same else "DIFF ")+label, "" if same else f"\n   ref={a}\n   new={b}")
rd=RD.from_size(20,1,1,comps); pd=P.Domain.from_size(20,1,1,comps)
rs=RS(rd,RAD(0.2,0.001,"FVM"),ics,bcs,rdefault("FVM")); ps=P.Simulation(pd,PAD(0.2,0.001,"FVM"),ics,bcs,P.default_scheme("FVM"))
both(lambda: np.array(rd.positions()), lambda: np.array(pd.positions()), "positions")
both(lambda: np.array(rd.values()), lambda: np.array(pd.values()), "values after ICs")
for split,locs in ((False,None),(True,[1]),(True,[1,2]),(True,[2])):
    both(lambda: np.array(rd.listify_interior(split,locs)), lambda: np.array(pd.listify_interior(split,locs)), f"listify {split} {locs}")
    b=[[-1.,0.,-2.],[5.,6.,7.]]
    both(lambda: rs.extend_bounds(b,20,3,split,locs), lambda: ps.extend_bounds(b,20,3,split,locs), f"extend_bounds {split} {locs}")
both(lambda: rs.extend_bounds(None,20,3), lambda: ps.extend_bounds(None,20,3), "extend_bounds None")
both(lambda: rs.extend_bounds([[1,2,3]],20,3), lambda: ps.extend_bounds([[1,2,3]],20,3), "extend_bounds len1")
both(lambda: rs.extend_bounds([[1,2],[1,2]],20,3), lambda: ps.extend_bounds([[1,2],[1,2]],20,3), "extend_bounds wrong nv")
both(lambda: rs.extend_bounds([[1,2,3],[1,2,3]],20,3,True,None), lambda: ps.extend_bounds([[1,2,3],[1,2,3]],20,3,True,None), "extend_bounds split no locs")
both(lambda: rs.extend_bounds([[1,2,3],[1,2,3]],20,3,True,[5]), lambda: ps.extend_bounds([[1,2,3],[1,2,3]],20,3,True,[5]), "extend_bounds bad locs")
both(lambda: rs.get_shape_from_list(list(range(60))), lambda: ps.get_shape_from_list(list(range(60))), "shape 60")
both(lambda: rs.get_shape_from_list(list(range(61))), lambda: ps.get_shape_from_list(list(range(61))), "shape 61")
both(lambda: (rd.ilo(),rd.ihi(),rd.num_components(),rd.component_index("v"),rd.component_name(2)), lambda: (pd.ilo(),pd.ihi(),pd.num_components(),pd.component_index("v"),pd.component_name(2)), "accessors")
both(lambda: [c.x() for c in rd.interior()][:5], lambda: [c.x() for c in pd.interior()][:5], "interior x")
both(lambda: [list(c.values()) for c in rd.boundaries()[0]+rd.boundaries()[1]], lambda: [list(c.values()) for c in pd.boundaries()[0]+pd.boundaries()[1]], "boundary values")
both(lambda: rs.initialize_from_list(list(np.arange(60.)),True,[1,2]) or np.array(rd.values()), lambda: ps.initialize_from_list(list(np.arange(60.)),True,[1,2]) or np.array(pd.values()), "initialize_from_list split")
both(lambda: rs.initialize_from_list(list(np.arange(59.)),False,None), lambda: ps.initialize_from_list(list(np.arange(59.)),False,None), "initialize_from_list bad len")
both(lambda: rs.initialize_from_list(list(np.arange(60.)),True,None), lambda: ps.initialize_from_list(list(np.arange(60.)),True,None), "initialize_from_list split no locs")
both(lambda: rs.initialize_from_list(list(np.arange(60.)),True,[4]), lambda: ps.initialize_from_list(list(np.arange(60.)),True,[4]), "initialize_from_list bad locs")
print("---- more")
from splitfxm.cell import Cell as RC
from splitfxm.boundary import Boundary as RB
from splitfxm.constants import btype as Rbt
from splitfxm.bc import apply_BC as rapply
from splitfxm.initialize import set_initial_condition as rset
from splitfxm.refine import Refiner as RRef
import copy
PC, PB, Pbt = P.Cell, P.Boundary, P.constants.btype if hasattr(P,'constants') else None
from splitfxm_b200.constants import btype as Pbt
from splitfxm_b200.bc import apply_BC as papply
from splitfxm_b200.initialize import set_initial_condition as pset
both(lambda: (RC(0.3,np.array([1.,2.]))==RC(0.3,np.array([5.,6.])), RC(0.1,np.array([1.]))<RC(0.2,np.array([1.]))), lambda: (PC(0.3,np.array([1.,2.]))==PC(0.3,np.array([5.,6.])), PC(0.1,np.array([1.]))<PC(0.2,np.array([1.]))), "Cell eq/lt")
def cellops(C):
    c=C(0.5,np.array([1.,2.,3.])); c.set_value(1,9.0); c.set_x(0.7); return (c.x(), list(c.values()), c.value(2))
both(lambda: cellops(RC), lambda: cellops(PC), "Cell set/get")
for args in ((-0.1,Rbt.LEFT,np.array([1.]),0.,1.),(0.5,Rbt.LEFT,np.array([1.]),0.,1.),(1.2,Rbt.LEFT,np.array([1.]),0.,1.),(1.2,Rbt.RIGHT,np.array([1.]),0.,1.),(-0.2,Rbt.RIGHT,np.array([1.]),0.,1.)):
    pargs=(args[0], Pbt.LEFT if args[1]==Rbt.LEFT else Pbt.RIGHT)+args[2:]
    both(lambda: (lambda b:(b.x(),list(b.values())))(RB(*args)), lambda: (lambda b:(b.x(),list(b.values())))(PB(*pargs)), f"Boundary{args[:2]}")
def setx(B,bt,x0,x1):
    b=B(x0,bt,np.array([1.]),0.,1.); b.set_x(x1); return b.x()
both(lambda: setx(RB,Rbt.LEFT,-0.1,0.5), lambda: setx(PB,Pbt.LEFT,-0.1,0.5), "Boundary.set_x interior")
both(lambda: setx(RB,Rbt.LEFT,-0.1,-0.3), lambda: setx(PB,Pbt.LEFT,-0.1,-0.3), "Boundary.set_x ok")
for ic in ("gaussian","rarefaction","tophat","sine","nope"):
    def run(Dm, setter):
        d=Dm.from_size(11,1,1,["a"]); setter(d,"a",ic); return np.array(d.values())
    both(lambda: run(RD,rset), lambda: run(P.Domain,pset), f"IC {ic}")
# apply_BC on object domains, every type, both sides, nb=2
for bc in ({"left":"periodic","right":"periodic"},{"left":"outflow","right":"outflow"},{"left":{"neumann":0.5},"right":{"neumann":-1.0}},{"left":{"dirichlet":2.0},"right":{"dirichlet":3.0}},{"left":"bogus","right":"periodic"},{"left":{"robin":1},"right":"periodic"},{"up":"periodic","right":"periodic"}):
    def run(Dm, setter, app):
        d=Dm.from_size(9,2,2,["a","b"]); setter(d,"a","gaussian"); setter(d,"b","sine"); app(d,"b",bc); return (np.array(d.positions()), np.array(d.values()))
    def cmp(f):
        x,v=f(); return np.concatenate([x, v.reshape(-1)])
    both(lambda: cmp(lambda: run(RD,rset,rapply)), lambda: cmp(lambda: run(P.Domain,pset,papply)), f"apply_BC {bc}")
both(lambda: (lambda d:(np.array(d.positions()),))(copy.deepcopy(rd))[0], lambda: (lambda d:(np.array(d.positions()),))(copy.deepcopy(pd))[0], "deepcopy")
for crit in ((1.5,.5,.5,0.),(3.,1.5,.5,0.),(3.,.5,-.1,0.),(3.,.5,.5,.6)):
    both(lambda: RRef().set_criteria(*crit), lambda: P.Refiner().set_criteria(*crit), f"set_criteria {crit}")
