"""audio_processing_tools_b200: B200-native hot path of Arable/audio_processing_tools.

Public surface mirrors the reference package for this path:
    processors.BaseProcessor / RainProcessor / has_processor
    noise_processor.NoiseProcessor
    edge.rain_signal_processor.SpectralNoiseProcessor / RainDetectorProcessor /
        NoiseProcessorConfig / build_noise_config
    audio_processing_framework.process_audio_batches_v2
"""
__version__ = "0.1.0"
