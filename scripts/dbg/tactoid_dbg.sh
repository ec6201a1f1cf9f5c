set -x
D=gpurun_out/tac; mkdir -p $D; cp oracle/_ref/examples/tactoid/disk.mesh $D/
python - <<'PY'
src=open('oracle/_ref/examples/tactoid/tactoid.morpho').read()
src=src[:src.index("// Function to visualize a director field")].replace("import plot\n","").replace("import povray\n","")
src+='\nprint "final ${m.count()} ${m.count(2)}"\nfor (en in problem.energies) print sopt.total(en).format("%.15g")\n'
open('gpurun_out/tac/stock.morpho','w').write(src)
open('gpurun_out/tac/ext.morpho','w').write("import morphocuda\n"+src+"print cudastats()\n")
PY
export LD_LIBRARY_PATH=$(cat oracle/_ref/openblas_dir.txt) OPENBLAS_NUM_THREADS=1 HOME=$PWD/oracle/_ref/home
cd $D
../../oracle/_ref/morpho6 stock.morpho > stock.out 2>&1
../../oracle/_ref/morpho6 ext.morpho > ext.out 2>&1
MORPHO_CUDA_PARANOID=1 ../../oracle/_ref/morpho6 ext.morpho > ext_paranoid.out 2>&1
wc -l *.out
diff stock.out ext.out | head -20
diff stock.out ext_paranoid.out | head -20
