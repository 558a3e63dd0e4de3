"""CPU oracle of the FF-LINS LiDAR hot path — TEST INFRASTRUCTURE ONLY (see oracle.h)."""
