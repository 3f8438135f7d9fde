class MixtureDensityNetwork:  # off the MPPI path; import-only stand-in
    pass
