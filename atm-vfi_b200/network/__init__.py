"""Drop-in counterparts of the reference's ``network`` package (network_base, network_lite, attention, flow_warp)."""
