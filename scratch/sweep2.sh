export GTA_B200_LPF_THREADS=448
for s in 1024 1152; do echo slots $s; GTA_B200_LPF_SLOTS=$s python scratch/lpf_time.py 100000 2>&1 | tail -2 | head -1; done
export GTA_B200_LPF_SLOTS=1152
echo "== cfg3 lpf"; GTA_B200_LPF=1 python scratch/lpf_vs_cta.py 3 8192,16384,32768,65536 2>&1 | tail -4
echo "== cfg3 fused"; GTA_B200_LPF=0 python scratch/lpf_vs_cta.py 3 2048,8192,16384,32768,65536 2>&1 | tail -5
echo "== cfg1"; GTA_B200_LPF=1 python scratch/lpf_vs_cta.py 1 32768,100000 2>&1 | tail -2; GTA_B200_LPF=0 python scratch/lpf_vs_cta.py 1 32768,100000 2>&1 | tail -2
echo "== cfg5"; GTA_B200_LPF=1 python scratch/lpf_vs_cta.py 5 10000,50000 2>&1 | tail -2; GTA_B200_LPF=0 python scratch/lpf_vs_cta.py 5 10000,50000 2>&1 | tail -2
