for cfgx in "12000 8000" "6000 4000" "3000 2000" "1500 1000"; do
  set -- $cfgx; wc=$1; mc=$2
  echo "== wave_cycles $wc min_cycles $mc"
  for cfg in "diabetes 2048" "munin 256"; do set -- $cfg
    PGX_MM_WAVE_CYCLES=$wc PGX_MM_MIN_CYCLES=$mc timeout 200 python tools/launch_profile.py $1 $2 1 2>&1 | sed -n 2p
  done
done
