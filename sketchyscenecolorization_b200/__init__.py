"""B200-native foreground-instance-colorization hot path (SketchySceneColorization drop-in)."""
__version__ = "0.1.0"
