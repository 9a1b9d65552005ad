# round-2 (z): ncu rows of the BatchNorm kernels changed after the r2h evidence set (OutConv-backward pair with eight
# loads in flight and no dy operand; strided kernels with eight pixels in flight).  Plain runs first.
set -x
S=gpurun_out
python tools/prof_step.py > $S/plain_step_r2z.log 2>&1 &&
timeout 300 ncu --profile-from-start off --set full --clock-control none -k 'regex:bn_bwd_.*contig|head_fwd' -c 4 -o $S/prof_head_r2z -f python tools/prof_step.py > $S/ncu_r2z_1.log 2>&1
python tools/prof_step.py SpectralUNET > $S/plain_spectral_r2z.log 2>&1 &&
timeout 300 ncu --profile-from-start off --set full --clock-control none -k 'regex:bn_' -s 9 -c 4 -o $S/prof_specbn_r2z -f python tools/prof_step.py SpectralUNET > $S/ncu_r2z_2.log 2>&1
python tools/ncu_summary.py $S/ncu_full_r2z_bn.csv $S/prof_head_r2z.ncu-rep $S/prof_specbn_r2z.ncu-rep && rm -f $S/prof_head_r2z.ncu-rep $S/prof_specbn_r2z.ncu-rep
cat $S/ncu_full_r2z_bn.csv | cut -c1-400
