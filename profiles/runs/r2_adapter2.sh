cd tests/golden/ref_inputs
B=../../../host/_ref/BEEP_b200
echo "== adapter England abcsmc"; BEEP_VERBOSE=1 timeout -s KILL 150 $B "inputfile=COVID 19 England/England_AS.toml" mode=abcsmc nsample=2000 ngeneration=3 cutoff_frac=0.5 outputdir=/tmp/o3 2>&1 | grep -v "^$\|Loaded" | tail -25
echo "== adapter EX_A1 filter"; timeout -s KILL 60 $B inputfile=examples/EX_A1.toml mode=pmcmc nparticle=256 nsample=10 outputdir=/tmp/o4 2>&1 | grep -v "^$\|Loaded" | tail -6
echo "== adapter England filter"; timeout -s KILL 100 $B "inputfile=COVID 19 England/England_AS.toml" mode=pmcmc nparticle=4096 nsample=10 outputdir=/tmp/o5 2>&1 | grep -v "^$\|Loaded" | tail -6
