#!/bin/bash
# round-end sequence on one GPU (tests, smoke, both bench arms) followed by the profile capture
bash scripts/gpu_full.sh
bash scripts/gpu_profile.sh
