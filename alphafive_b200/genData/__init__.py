"""Drop-in counterparts of the reference's ``genData`` package (player, network, networkAPI)."""
