import os,sys
sys.path.insert(0,'/root/repo/tools')
import bench_decode as bd
for shape in ((4096,11,64,64),(16384,17,96,72),(2048,11,128,128),(512,11,64,64),(64,11,64,64)):
    bd.run(*shape)
