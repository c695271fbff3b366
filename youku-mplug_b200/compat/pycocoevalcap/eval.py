class COCOEvalCap:
    def __init__(self, *args, **kwargs):
        raise ImportError("caption metrics need the real `pycocoevalcap` package (pip install pycocoevalcap); "
                          "this stand-in only satisfies the script's top-level import")
