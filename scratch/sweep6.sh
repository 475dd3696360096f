for cv in 44 58 72; do echo "carve $cv: $(GTA_B200_LPF_CARVE=$cv python scratch/lpf_vs_cta.py 3 66000 | tail -1)"; done
for cv in 58 72; do echo "carve $cv: $(GTA_B200_LPF_CARVE=$cv python scratch/lpf_vs_cta.py 3 100000 | tail -1)"; done
