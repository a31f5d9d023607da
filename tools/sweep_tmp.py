import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import quickbench as q
N=(256,256,128)
q.variant('tma','tma')
q.run(*N,1,1,tag='default')
q.run(*N,1,0,tag='o2 inviscid')
q.run(*N,0,0,tag='o1 inviscid')
q.run(256,128,64,1,1,ptype=1,lx=2.0,ly=0.008,lz=1.0,dt=3e-8,tag='flatplate')
