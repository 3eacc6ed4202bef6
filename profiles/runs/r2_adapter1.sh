cd tests/golden/ref_inputs
B=../../../host/_ref/BEEP_b200; R=../../../oracle/_ref/BEEP
echo "== adapter EX_A1 abcsmc"; time $B inputfile=examples/EX_A1.toml mode=abcsmc nsample=2000 ngeneration=6 cutoff_frac=0.5 outputdir=/tmp/o1 2>&1 | grep -v "^$" | tail -12
echo "== reference EX_A1 abcsmc"; time $R inputfile=examples/EX_A1.toml mode=abcsmc nsample=2000 ngeneration=6 cutoff_frac=0.5 outputdir=/tmp/o2 2>&1 | grep -i "generation\|cut" | tail -12
echo "== adapter England abcsmc"; time $B "inputfile=COVID 19 England/England_AS.toml" mode=abcsmc nsample=1000 ngeneration=3 cutoff_frac=0.5 outputdir=/tmp/o3 2>&1 | grep -v "^$" | tail -8
echo "== adapter EX_A1 filter"; $B inputfile=examples/EX_A1.toml mode=pmcmc nparticle=256 nsample=10 outputdir=/tmp/o4 2>&1 | grep -v "^$" | tail -6
echo "== adapter England filter"; $B "inputfile=COVID 19 England/England_AS.toml" mode=pmcmc nparticle=4096 nsample=10 outputdir=/tmp/o5 2>&1 | grep -v "^$" | tail -6
