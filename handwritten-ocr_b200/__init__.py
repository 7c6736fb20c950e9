"""handwritten-ocr_b200: B200-native OCR read path (preprocess -> VLM read -> agreement/merge/CER)
behind the reference's `ocr_agent.tools` surface.  See DESIGN.md."""
__version__ = "0.1.0"
