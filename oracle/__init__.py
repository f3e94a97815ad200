"""CPU oracle for the HiFiGAN hot path.  Test infrastructure only -- see hifigan_oracle.py."""
