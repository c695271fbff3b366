"""Import-time stand-in for pycocoevalcap (see compat/pycocotools/__init__.py): the reference's caption script imports
`pycocoevalcap.eval.COCOEvalCap` at module level (downstream/run_caption_distributed_gpt3.py:40) and uses it only to
score generated captions (:293-296)."""
