"""yoxel-voxel_b200 — B200-native SVO ray caster (host-side mirror of the reference renderer API).

Import as ``yoxel_voxel_b200`` (see the shim package next to this directory)."""
