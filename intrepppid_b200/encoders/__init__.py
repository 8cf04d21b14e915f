from .awd_lstm import AWDLSTM, AWDLSTMEncoder, Projection

__all__ = ["AWDLSTM", "AWDLSTMEncoder", "Projection"]
