class Artist:  # marker only
    pass
