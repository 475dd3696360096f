for st in 704 768 1024; do for cv in 0 60 75 90 100; do echo "stride $st carve $cv: $(GTA_B200_LPF_STRIDE=$st GTA_B200_LPF_CARVE=$cv python scratch/lpf_vs_cta.py 3 100000 | tail -1)"; done; done
