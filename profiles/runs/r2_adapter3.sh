cd tests/golden/ref_inputs
B=../../../host/_ref/BEEP_b200
echo "== adapter England abcsmc prop_size 0.05"; BEEP_VERBOSE=1 timeout -s KILL 200 $B "inputfile=COVID 19 England/England_AS.toml" mode=abcsmc nsample=8192 ngeneration=4 cutoff_frac=0.5 prop_size=0.05 outputdir=/tmp/o3 2>&1 | grep "generation\|ERROR\|rror" | tail -25
echo "== adapter England filter"; BEEP_RUN=filter BEEP_NPARTICLE=4096 timeout -s KILL 100 $B "inputfile=COVID 19 England/England_AS.toml" mode=abcsmc nsample=10 ngeneration=2 outputdir=/tmp/o5 2>&1 | grep "filter\|ERROR\|rror" | tail -6
