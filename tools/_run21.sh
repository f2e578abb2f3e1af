for cfgx in "512 4" "256 4" "1024 4" "512 2" "512 8" "256 8" "1024 2"; do
  set -- $cfgx; to=$1; btb=$2
  echo "== TO cap $to btb $btb"
  for cfg in "diabetes 2048" "munin 256" "pathfinder 16384"; do set -- $cfg
    PGX_TILE_TO_CAP=$to PGX_TILE_BTB=$btb timeout 200 python tools/launch_profile.py $1 $2 1 2>&1 | sed -n 2p
  done
done
