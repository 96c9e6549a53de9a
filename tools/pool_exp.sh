#!/bin/bash
# pooling epilogue: cluster size (token slices per sequence) x ring depth on the experiments build
# (make EXPERIMENTS=1 OBJDIR=build_x OUT=../libprs_x.so; EXTRA=-DPOOL_PER=8 for 8 rows per thread and stage);
# libprs_old.so = the previous kernel when present
for shape in "256 512 768" "64 256 384" "1024 128 768"; do for m in full random; do for dt in float16 float32; do
  [ -f persian-rag-system_b200/libprs_old.so ] && PRS_LIB_PATH=$PWD/persian-rag-system_b200/libprs_old.so python tools/prof_pool.py $shape $dt $m | sed 's/^/OLD /'
  python tools/prof_pool.py $shape $dt $m | sed 's/^/NEW /'
done; done; done
export PRS_LIB_PATH=$PWD/persian-rag-system_b200/libprs_x.so
python tools/prof_pool.py 256 512 768 float16 full sweep
python tools/prof_pool.py 256 512 768 float16 random sweep
