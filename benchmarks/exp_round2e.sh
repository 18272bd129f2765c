set -u
bash benchmarks/round_check.sh
