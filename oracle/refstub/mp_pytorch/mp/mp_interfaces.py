class MPInterface:  # type stub only: the real arithmetic lives in mp_pytorch (absent here)
    pass
